"""A randomised-interleaving model of the persistent stream kernel's protocol (sema_b200/csrc/k2_stream.cuh).

No GPU and no product code: this is the ORDERING ARGUMENT of the kernel header restated as a small state machine and
run under thousands of random schedules, so that the invariants the kernel relies on are checked rather than argued:

  * a consumer parks query q's lists in shared-memory slot q & 1 only after the finisher has read that slot's
    previous use (slot_free), and the finisher never reads a slot before all of q's lists are in it (parked);
  * a finisher posts query q (block list of parity q & 1, ticket of parity q & 1) only once q - 2 is complete
    (done >= q - 1): the merge of q - 2 never sees a list of q;
  * queries complete in order on every rank (done goes 0, 1, 2, ...), because a block's finisher is sequential;
  * a rank publishes query q to its peers only after it has consumed q - 1, so the two alternating exchange slots are
    never overwritten while a peer still reads them, although ranks run at different speeds.

Each actor (a block's scan side, a block's finisher, per rank) advances one small step at a time; multi-step reads
(the last block's merge over all block lists, the merge of the exchanged keys) re-check what they read at their end.
The same model with the flow-control wait removed must FAIL — that is what makes the passing runs meaningful.
"""
import random

import pytest


class Violation(AssertionError):
    pass


class Rank:
    def __init__(self, world, rank, blocks, nq):
        self.world, self.rank, self.B, self.nq = world, rank, blocks, nq
        self.done = 0
        self.ticket = [0, 0]
        self.partials = [[None] * blocks for _ in range(2)]          # query whose list sits in [parity][block]
        self.xkeys = [[None] * world for _ in range(2)]              # this rank's exchange buffer: [slot][source] = seq
        self.xflag = [[-1] * world for _ in range(2)]
        self.results = []
        # per block: scan side and finisher
        self.scan_q = [0] * blocks                                   # query the consumers are scanning
        self.slot = [[None, None] for _ in range(blocks)]            # parked lists: [block][parity] = query
        self.slot_free = [[True, True] for _ in range(blocks)]
        self.fin_q = [0] * blocks
        self.fin_phase = ["wait_parked"] * blocks                     # wait_parked -> wait_done -> (merge -> publish -> wait_peers -> xmerge) -> ...
        self.fin_left = [0] * blocks                                  # remaining steps of a multi-step read


def runnable(ranks, flow_control):
    acts = []
    for r in ranks:
        for b in range(r.B):
            q = r.scan_q[b]
            if q < r.nq and r.slot_free[b][q & 1]:
                acts.append((r, b, "scan"))
            f, ph = r.fin_q[b], r.fin_phase[b]
            if f >= r.nq:
                continue
            if ph == "wait_parked" and r.slot[b][f & 1] == f:
                acts.append((r, b, "fin"))
            elif ph == "wait_done" and (not flow_control or f < 2 or r.done >= f - 1):
                acts.append((r, b, "fin"))
            elif ph in ("merge", "publish", "xmerge"):
                acts.append((r, b, "fin"))
            elif ph == "wait_peers" and all(x >= f for x in r.xflag[f & 1]):
                acts.append((r, b, "fin"))
    return acts


def step(ranks, r, b, kind, rng):
    if kind == "scan":                       # the consumers finish scanning query q on this block and park their lists
        q = r.scan_q[b]
        if r.slot[b][q & 1] is not None and not r.slot_free[b][q & 1]:
            raise Violation("parked over lists the finisher has not read")
        r.slot[b][q & 1] = q
        r.slot_free[b][q & 1] = False
        r.scan_q[b] = q + 1
        return
    f, ph = r.fin_q[b], r.fin_phase[b]
    p = f & 1
    if ph == "wait_parked":                  # block merge of the parked lists, slot handed back
        if r.slot[b][p] != f:
            raise Violation("finisher read a slot that does not hold its query")
        r.slot_free[b][p] = True
        r.fin_phase[b] = "wait_done"
    elif ph == "wait_done":                  # post: block list + ticket of parity p
        r.partials[p][b] = f
        r.ticket[p] += 1
        if r.ticket[p] == r.B:
            r.fin_phase[b] = "merge"
            r.fin_left[b] = rng.randint(1, 6)
        else:
            r.fin_q[b] = f + 1
            r.fin_phase[b] = "wait_parked"
    elif ph == "merge":                      # last block: multi-step read of every block list of parity p
        r.fin_left[b] -= 1
        if r.fin_left[b] > 0:
            return
        if any(x != f for x in r.partials[p]):
            raise Violation(f"merge of query {f} saw block lists {r.partials[p]}")
        r.fin_phase[b] = "publish"
    elif ph == "publish":                    # fused exchange: keys into every rank's slot p, then the flags
        if r.done != f:
            raise Violation(f"rank {r.rank} publishes query {f} with done = {r.done}")
        for g in ranks:
            g.xkeys[p][r.rank] = f
            g.xflag[p][r.rank] = f
        r.fin_phase[b] = "wait_peers"
    elif ph == "wait_peers":
        r.fin_phase[b] = "xmerge"
        r.fin_left[b] = rng.randint(1, 4)
    elif ph == "xmerge":                     # multi-step read of the world x k exchanged keys
        r.fin_left[b] -= 1
        if r.fin_left[b] > 0:
            return
        if any(x != f for x in r.xkeys[p]):
            raise Violation(f"rank {r.rank} merged query {f} from exchange slots {r.xkeys[p]}")
        r.ticket[p] = 0
        if r.done != f:
            raise Violation("queries completed out of order")
        r.done = f + 1
        r.results.append(f)
        r.fin_q[b] = f + 1
        r.fin_phase[b] = "wait_parked"


def simulate(seed, world=2, blocks=4, nq=10, flow_control=True, bias=None):
    rng = random.Random(seed)
    ranks = [Rank(world, g, blocks, nq) for g in range(world)]
    for _ in range(200000):
        acts = runnable(ranks, flow_control)
        if not acts:
            break
        if bias is not None:                 # starve one actor class most of the time: skewed speeds
            pref = [a for a in acts if bias(a)]
            if pref and rng.random() < 0.9:
                acts = pref
        r, b, kind = rng.choice(acts)
        step(ranks, r, b, kind, rng)
    for r in ranks:
        if r.results != list(range(nq)):
            raise Violation(f"rank {r.rank} finished {r.results} (deadlock or reordering)")
    return True


@pytest.mark.parametrize("world,blocks,nq", [(1, 1, 7), (1, 3, 9), (2, 4, 10), (3, 2, 8), (4, 3, 6)])
def test_protocol_holds_under_random_schedules(world, blocks, nq):
    for seed in range(150):
        assert simulate(seed, world, blocks, nq)


def test_protocol_holds_with_skewed_speeds():
    # scan far ahead of the finishers; one rank far ahead of the others; one block's finisher always last
    biases = [lambda a: a[2] == "scan", lambda a: a[2] == "fin", lambda a: a[0].rank == 0, lambda a: a[1] != 0,
              lambda a: a[0].rank != 0 and a[2] == "scan"]
    for i, bias in enumerate(biases):
        for seed in range(60):
            assert simulate(1000 * i + seed, 3, 4, 9, bias=bias)


def test_model_detects_the_race_the_flow_control_prevents():
    """Without `done >= q - 1` before posting, a fast block overwrites the list of q - 2 under the merge: the model
    must notice, otherwise the tests above prove nothing."""
    hits = 0
    for seed in range(300):
        try:
            simulate(seed, 1, 4, 10, flow_control=False, bias=lambda a: a[2] != "fin" or a[0].fin_phase[a[1]] != "merge")
        except Violation:
            hits += 1
    assert hits > 0
