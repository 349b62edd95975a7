// prints how many clusters of each size can be resident with a 1-CTA/SM kernel (200 KB smem, 192 threads)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(192, 1) k(int *p) { extern __shared__ char s[]; if (p) p[0] = s[0]; }
int main()
{
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    printf("%s SMs=%d\n", pr.name, pr.multiProcessorCount);
    for (int c = 1; c <= 16; ++c) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(c, 148, 1); cfg.blockDim = dim3(192, 1, 1); cfg.dynamicSmemBytes = 200 * 1024;
        cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim.x = c; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
        cfg.attrs = a; cfg.numAttrs = 1;
        int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
        printf("cluster %2d: max active clusters %3d -> %3d SMs (%s)\n", c, n, n * c, cudaGetErrorString(e));
    }
    return 0;
}
