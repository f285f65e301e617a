#include <cstdio>
#include <cstdint>
__global__ void k(int n) {
  __shared__ uint64_t bars[20];
  __shared__ float stats[512];
  __shared__ uint32_t slot[4];
  extern __shared__ __align__(1024) uint8_t dyn[];
  if (threadIdx.x == 0) {
    bars[0] = n; stats[0] = n; slot[0] = n;
    printf("bars %u stats %u slot %u dyn %u (dyn %% 1024 = %u)\n", (unsigned)__cvta_generic_to_shared(bars), (unsigned)__cvta_generic_to_shared(stats),
           (unsigned)__cvta_generic_to_shared(slot), (unsigned)__cvta_generic_to_shared(dyn), (unsigned)__cvta_generic_to_shared(dyn) % 1024);
    dyn[n] = 1;
  }
}
__global__ void k2(int n) {
  extern __shared__ __align__(1024) uint8_t dyn[];
  if (threadIdx.x == 0) { printf("no static: dyn %u\n", (unsigned)__cvta_generic_to_shared(dyn)); dyn[n] = 1; }
}
int main() {
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 229376);
  k<<<1, 32, 229376>>>(5); printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  k<<<1, 32, 4096>>>(5); cudaDeviceSynchronize();
  cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  k2<<<1, 32, 232448>>>(5); printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
