// stand-alone probe of the TMA staging used by vsl_fused.cu (debug aid, not part of the library)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
struct alignas(64) Args { int W, H; int pad[14]; CUtensorMap tm; };
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ Args a, float* out, int x, int y, int z, int bytes) {
  extern __shared__ __align__(128) unsigned char raw[];
  float* tile = reinterpret_cast<float*>(raw);
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(raw + 3 * 20 * 36 * 4 + 64);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                     smem_u32(tile)),
                 "l"(reinterpret_cast<unsigned long long>(&a.tm)), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
                 : "memory");
  }
  asm volatile(
      "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(bar)),
      "r"(0)
      : "memory");
  for (int i = threadIdx.x; i < 3 * 20 * 36; i += blockDim.x) out[i] = tile[i];
}
int main(int argc, char** argv) {
  const int BW = argc > 1 ? atoi(argv[1]) : 36, X = argc > 2 ? atoi(argv[2]) : 30, NOSYNC = argc > 3 ? atoi(argv[3]) : 0;
  (void)NOSYNC;
  const int B = 2, H = 64, W = 96;
  std::vector<float> h((size_t)B * 3 * H * W);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
  float *d, *o;
  cudaMalloc(&d, h.size() * 4);
  cudaMalloc(&o, 3 * 20 * 36 * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
  printf("entry point: %d %d %p\n", (int)e, (int)q, sym);
  typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                         const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  Args a = {};
  a.W = W; a.H = H;
  const cuuint64_t dims[3] = {W, H, B * 3};
  const cuuint64_t strides[2] = {W * 4ull, (cuuint64_t)W * H * 4};
  const cuuint32_t box[3] = {(cuuint32_t)BW, 20, 3};
  const cuuint32_t es[3] = {1, 1, 1};
  CUresult r = ((Fn)sym)(&a.tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d\n", (int)r);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
  probe<<<1, 128, 60000>>>(a, o, X, 14, 3, 3 * 20 * BW * 4);
  e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  std::vector<float> t(3 * 20 * 36);
  cudaMemcpy(t.data(), o, t.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int c = 0; c < 3; ++c) for (int i = 0; i < 20; ++i) for (int j = 0; j < 36; ++j) {
    if (j >= BW) continue;
    float want = (float)(((size_t)(3 + c) * H + 14 + i) * W + X + j);
    if (t[(c * 20 + i) * BW + j] != want) ++bad;
  }
  printf("mismatches: %d  first %f\n", bad, t[0]);
  return 0;
}
