#!/bin/bash
# Opcode evidence for the shipped library: per kernel, how many tcgen05 / TMA / mbarrier SASS instructions it holds.
# Usage: scripts/sass_histogram.sh > profiles/r02_sass_opcodes.txt   (CPU only: reads sema_b200/libsema_b200.so)
SO=${1:-sema_b200/libsema_b200.so}
echo "# cuobjdump -sass $SO  ($(date -u +%F), $(/usr/local/cuda/bin/nvcc --version | tail -1))"
echo "# columns: kernel | UTCHMMA (tcgen05.mma) [of which .2CTA] | UTCBAR (tcgen05.commit) | LDTM/STTM (tcgen05.ld/st) | UBLKCP (cp.async.bulk) [of which MULTICAST] | UTMALDG (cp.async.bulk.tensor) | SYNCS (mbarrier)"
cuobjdump -sass "$SO" | awk '
/Function : /{ if (name != "") out(); name=$3; delete c }
/UTCHMMA/{c["mma"]++} /UTCHMMA.2CTA/{c["mma2"]++} /UTCBAR/{c["bar"]++} /LDTM|STTM/{c["tm"]++}
/UBLKCP/{c["blk"]++} /UBLKCP.*MULTICAST/{c["blkm"]++} /UTMALDG/{c["tma"]++} /SYNCS/{c["syncs"]++}
function out(){ if (c["mma"]+c["blk"]+c["tma"]+c["tm"] > 0) printf "%s | %d [%d] | %d | %d | %d [%d] | %d | %d\n", name, c["mma"], c["mma2"], c["bar"], c["tm"], c["blk"], c["blkm"], c["tma"], c["syncs"] }
END{ out() }' | while IFS= read -r line; do n=$(echo "$line" | cut -d' ' -f1 | c++filt); echo "$n |${line#*|}"; done
