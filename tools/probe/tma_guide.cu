// The CUDA programming guide's TMA tensor-tile example (libcu++ wrappers), as a sanity probe of the box.
#include <cuda.h>
#include <cuda/barrier>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
constexpr int GW = 1024, GH = 1024, SW = 32, SH = 8;
__global__ void kernel(const __grid_constant__ CUtensorMap tensor_map, int x, int y, int* out) {
  __shared__ alignas(128) int smem_buffer[SH][SW];
#pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ barrier bar;
  if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
  __syncthreads();
  barrier::arrival_token token;
  if (threadIdx.x == 0) {
    cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
    token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
  } else {
    token = bar.arrive();
  }
  bar.wait(std::move(token));
  if (threadIdx.x == 0) { out[0] = smem_buffer[0][0]; out[1] = smem_buffer[SH - 1][SW - 1]; }
}
int main() {
  std::vector<int> h((size_t)GW * GH);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (int)i;
  int *d, *out;
  cudaMalloc(&d, h.size() * 4);
  cudaMalloc(&out, 8);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  auto encode = (CUresult(*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill))fn;
  CUtensorMap tm{};
  cuuint64_t size[2] = {GW, GH}, stride[1] = {GW * sizeof(int)};
  cuuint32_t box[2] = {SW, SH}, es[2] = {1, 1};
  CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, d, size, stride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                      CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  kernel<<<1, 128>>>(tm, 64, 16, out);
  cudaError_t e = cudaDeviceSynchronize();
  int ho[2] = {0, 0};
  cudaMemcpy(ho, out, 8, cudaMemcpyDeviceToHost);
  printf("guide example: encode rc=%d, %s, got %d %d want %d %d\n", (int)r, cudaGetErrorString(e), ho[0], ho[1], 16 * GW + 64, (16 + SH - 1) * GW + 64 + SW - 1);
  return 0;
}
