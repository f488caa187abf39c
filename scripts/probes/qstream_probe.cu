// Probe: how fast can a B200 stream the fp32 Q table (16 B/site, read once + written once per
// iteration) through per-warp TMA rings, as a function of the box shape, the ring depth and the
// store path?  No stencil work: this is the data-path ceiling of k_step_fast.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/qstream_probe scripts/probes/qstream_probe.cu
//   build/qstream_probe            (prints one line per configuration)
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done, spins = 0;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (!done && ++spins > (1u << 26)) __trap();
  } while (!done);
}
__device__ __forceinline__ uint64_t evict_first() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1, int c2, uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
               ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *m, const void *src, int c0, int c1, int c2, uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3, %4}], [%1], %5;"
               ::"l"(m), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "l"(pol) : "memory");
}

// Q as bytes: dims (128, L/8, rows); box (128, bw/128, br).  A job = one box; jobs are dealt to
// warps so that the `wpc` warps of a CTA take vertically adjacent boxes of one column block
// (like k_step_fast's tile) and CTAs walk the tile list with a stride of gridDim.x.
// STORE: 0 = TMA store from shared memory, 1 = st.global.v4 from registers (streaming)
template <int NB, int STORE>
__global__ void __launch_bounds__(256) k_probe(const __grid_constant__ CUtensorMap map, float4 *Q, int L, int rows,
                                                int bw, int br, int dist) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
  const int boxB = bw * br;
  unsigned char *ring = smem + (size_t)warp * NB * boxB;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)wpc * NB * boxB) + warp * NB;
  if (lane == 0) { for (int b = 0; b < NB; ++b) mbar_init(&bars[b], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncwarp();
  const uint64_t pol = evict_first();
  const int n_bx = (L * 16) / bw, n_by = rows / (br * wpc);   // CTA tiles: wpc boxes high
  const int n_tiles = n_bx * n_by;
  // job k of this warp: CTA tile t = blockIdx.x + k * gridDim.x
  auto coords = [&](int k, int &c1, int &c2) {
    const int t = blockIdx.x + k * gridDim.x;
    const int ty = t / n_bx, tx = t - ty * n_bx;
    c1 = tx * (bw / 128);
    c2 = (ty * wpc + warp) * br;
  };
  const int n_jobs = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto issue = [&](int k) {
    int c1, c2; coords(k, c1, c2);
    const int b = k % NB;
    mbar_expect_tx(&bars[b], boxB);
    tma_load_3d(ring + (size_t)b * boxB, &map, &bars[b], 0, c1, c2, pol);
  };
  if (lane == 0) for (int k = 0; k < dist && k < n_jobs; ++k) issue(k);
  for (int k = 0; k < n_jobs; ++k) {
    const int b = k % NB;
    if (lane == 0 && k + dist < n_jobs) {
      // the buffer of job k+dist was last used by job k+dist-NB: its store must have left smem
      if (STORE == 0) {
        if (NB - dist >= 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
      issue(k + dist);
    }
    mbar_wait(&bars[b], (uint32_t)(k / NB) & 1u);
    unsigned char *buf = ring + (size_t)b * boxB;
    int c1, c2; coords(k, c1, c2);
    // touch every float4 once (lane l: 16-byte chunks l, l+32, ...), like an in-place update
    for (int o = lane * 16; o < boxB; o += 512) {
      float4 v = *reinterpret_cast<float4 *>(buf + o);
      v.x += 1.0f;
      if (STORE == 0) *reinterpret_cast<float4 *>(buf + o) = v;
      else {
        // undo the 128-byte swizzle: chunk index within the 1024-byte atom
        const int line = o >> 7, ch = ((o >> 4) & 7) ^ (line & 7);
        const int row = line / (bw / 128), cl = line % (bw / 128);
        const size_t gb = ((size_t)(c2 + row) * L * 16) + (size_t)(c1 + cl) * 128 + ch * 16;
        __stcs(reinterpret_cast<float4 *>(reinterpret_cast<unsigned char *>(Q) + gb), v);
      }
    }
    if (STORE == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) { tma_store_3d(&map, buf, 0, c1, c2, pol); asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
    } else __syncwarp();
  }
  if (STORE == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// reference point: plain grid-stride float4 read-modify-write (no smem)
__global__ void k_rmw(float4 *Q, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float4 v = __ldcs(Q + i); v.x += 1.0f; __stcs(Q + i, v);
  }
}

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  void *fn = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  return (PFN_cuTensorMapEncodeTiled_v12000)fn;
}

template <int NB, int STORE>
static void run(float4 *Q, int L, int rows, int bw, int br, int wpc, int ctas_per_sm, int dist) {
  static auto encode = get_encode();
  CUtensorMap map;
  const cuuint64_t dims[3] = {128, (cuuint64_t)L / 8, (cuuint64_t)rows};
  const cuuint64_t strides[2] = {128, (cuuint64_t)L * 16};
  const cuuint32_t box[3] = {128, (cuuint32_t)(bw / 128), (cuuint32_t)br};
  const cuuint32_t es[3] = {1, 1, 1};
  if (encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, Q, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
    printf("encode failed\n"); return;
  }
  const size_t smem = (size_t)wpc * NB * bw * br + wpc * NB * 8 + 1024 + 64;
  if (smem > 227 * 1024 / ctas_per_sm - 1024) { printf("NB=%d store=%d box=%dx%d wpc=%d ctas/SM=%d: smem %zu too large\n", NB, STORE, bw, br, wpc, ctas_per_sm, smem); return; }
  CK(cudaFuncSetAttribute(k_probe<NB, STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = 148 * ctas_per_sm;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) k_probe<NB, STORE><<<grid, wpc * 32, smem>>>(map, Q, L, rows, bw, br, dist);
  CK(cudaDeviceSynchronize());
  const int reps = 20;
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) k_probe<NB, STORE><<<grid, wpc * 32, smem>>>(map, Q, L, rows, bw, br, dist);
  CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  const double us = 1e3 * ms / reps, gbs = 2.0 * L * (double)rows * 16 / (us * 1e-6) / 1e9;
  printf("NB=%d dist=%d store=%s box=%4dB x %d rows  warps/CTA=%2d CTAs/SM=%d  smem/CTA=%6zu : %7.2f us  %7.1f GB/s (r+w)\n", NB, dist,
         STORE ? "stg" : "tma", bw, br, wpc, ctas_per_sm, smem, us, gbs);
}

int main() {
  const int L = 4096, rows = 4096;
  float4 *Q; CK(cudaMalloc(&Q, (size_t)L * rows * 16)); CK(cudaMemset(Q, 0, (size_t)L * rows * 16));
  {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int g : {148 * 8, 148 * 16, 148 * 32}) {
      k_rmw<<<g, 256>>>(Q, (size_t)L * rows); CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0)); for (int i = 0; i < 20; ++i) k_rmw<<<g, 256>>>(Q, (size_t)L * rows); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      printf("plain float4 rmw grid=%d: %.2f us  %.1f GB/s (r+w)\n", g, 1e3 * ms / 20, 2.0 * L * (double)rows * 16 / (1e3 * ms / 20 * 1e-6) / 1e9);
    }
  }
  // the current kernel's shape: 8 warps x (2 KB x 2 rows), 2 buffers, 2 CTAs/SM
  run<2, 0>(Q, L, rows, 2048, 2, 8, 2, 1);
  run<3, 0>(Q, L, rows, 2048, 2, 8, 1, 2);
  run<4, 0>(Q, L, rows, 2048, 1, 8, 2, 2);
  run<4, 0>(Q, L, rows, 2048, 1, 8, 2, 3);
  run<2, 1>(Q, L, rows, 2048, 2, 8, 2, 1);
  run<3, 1>(Q, L, rows, 2048, 1, 8, 2, 2);
  run<4, 1>(Q, L, rows, 2048, 1, 8, 2, 3);
  run<4, 1>(Q, L, rows, 2048, 1, 8, 2, 2);
  run<4, 0>(Q, L, rows, 4096, 1, 4, 4, 2);
  run<2, 0>(Q, L, rows, 4096, 1, 8, 2, 1);
  run<2, 1>(Q, L, rows, 4096, 1, 8, 2, 1);
  run<4, 1>(Q, L, rows, 4096, 1, 4, 2, 3);
  run<2, 0>(Q, L, rows, 8192, 1, 4, 2, 1);
  run<2, 1>(Q, L, rows, 8192, 1, 4, 2, 1);
  run<6, 0>(Q, L, rows, 2048, 1, 8, 2, 4);
  run<6, 1>(Q, L, rows, 2048, 1, 8, 2, 5);
  run<3, 0>(Q, L, rows, 2048, 2, 8, 2, 2);   // does not fit with the real kernel's other planes; ceiling only
  run<3, 1>(Q, L, rows, 2048, 2, 8, 2, 2);
  run<2, 0>(Q, L, rows, 2048, 2, 16, 1, 1);
  run<2, 0>(Q, L, rows, 2048, 2, 4, 4, 1);
  return 0;
}
