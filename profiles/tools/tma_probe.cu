// Stand-alone probe of the TMA tile load used by k_track_features (u8 image, 48x48 box, one warp).
// nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu && ./tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define TILE 48
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int BOXW, int BOXH>
__global__ void k(const __grid_constant__ CUtensorMap tmap, const CUtensorMap* gmap, int use_global, int x0, int y0,
                  unsigned char* out) {
  __shared__ __align__(128) unsigned char tile[BOXW * BOXH];
  __shared__ __align__(8) unsigned long long bar;
  const CUtensorMap* m = use_global ? gmap : &tmap;
  if (threadIdx.x == 0) {
    unsigned b = smem_addr(&bar), d = smem_addr(tile);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(BOXW * BOXH) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(d),
                 "l"(m), "r"(x0), "r"(y0), "r"(b) : "memory");
  }
  __syncwarp();
  unsigned ok = 0, b = smem_addr(&bar);
  while (!ok)
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(b) : "memory");
  for (int i = threadIdx.x; i < BOXW * BOXH; i += 32) out[i] = tile[i];
}
template <int BOXW, int BOXH>
int run(EncodeTiledFn fn, unsigned char* d_img, int w, int h, int pitch, const std::vector<unsigned char>& img, int use_global) {
  CUtensorMap map;
  cuuint64_t dims[2] = {(cuuint64_t)w, (cuuint64_t)h};
  cuuint64_t strides[1] = {(cuuint64_t)pitch};
  cuuint32_t box[2] = {BOXW, BOXH};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d_img, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("box %dx%d global=%d encode rc=%d  ", BOXW, BOXH, use_global, (int)r);
  if (r != CUDA_SUCCESS) { printf("\n"); return 1; }
  CUtensorMap* gmap; cudaMalloc(&gmap, sizeof(map)); cudaMemcpy(gmap, &map, sizeof(map), cudaMemcpyHostToDevice);
  unsigned char* d_out; cudaMalloc(&d_out, BOXW * BOXH);
  const int x0 = getenv("X0") ? atoi(getenv("X0")) : 37, y0 = getenv("Y0") ? atoi(getenv("Y0")) : 21;
  k<BOXW, BOXH><<<1, 32>>>(map, gmap, use_global, x0, y0, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("kernel: %s\n", cudaGetErrorString(e)); return 2; }
  std::vector<unsigned char> o(BOXW * BOXH);
  cudaMemcpy(o.data(), d_out, o.size(), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int y = 0; y < BOXH; ++y) for (int x = 0; x < BOXW; ++x) bad += o[y * BOXW + x] != img[(size_t)(y0 + y) * pitch + x0 + x];
  printf("mismatches %d\n", bad);
  return bad != 0;
}
int main() {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  printf("entry %p q=%d\n", p, (int)q);
  EncodeTiledFn fn = (EncodeTiledFn)p;
  const int w = getenv("IW") ? atoi(getenv("IW")) : 650, h = 490, pitch = 704;
  std::vector<unsigned char> img((size_t)pitch * h);
  for (size_t i = 0; i < img.size(); ++i) img[i] = (unsigned char)(i * 2654435761u >> 24);
  unsigned char* d_img; cudaMalloc(&d_img, img.size()); cudaMemcpy(d_img, img.data(), img.size(), cudaMemcpyHostToDevice);
  int rc = 0;
  const char* only = getenv("PROBE");
  int sel = only ? atoi(only) : -1;
  if (sel < 0 || sel == 0) rc |= run<48, 48>(fn, d_img, w, h, pitch, img, 0);
  if (sel < 0 || sel == 1) rc |= run<48, 48>(fn, d_img, w, h, pitch, img, 1);
  if (sel < 0 || sel == 2) rc |= run<64, 32>(fn, d_img, w, h, pitch, img, 0);
  if (sel < 0 || sel == 3) rc |= run<64, 32>(fn, d_img, w, h, pitch, img, 1);
  if (sel < 0 || sel == 4) rc |= run<32, 48>(fn, d_img, w, h, pitch, img, 1);
  if (sel < 0 || sel == 5) rc |= run<16, 48>(fn, d_img, w, h, pitch, img, 1);
  return rc;
}
