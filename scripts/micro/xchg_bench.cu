// xchg_bench.cu -- what does a grid-wide exchange among ~128 co-resident blocks cost on this GPU?  (design input for
// cssm_series.cuh; not part of the library).  nvcc -O3 -arch=sm_100a -o xchg_bench xchg_bench.cu && ./xchg_bench
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
typedef unsigned long long u64;

__device__ __forceinline__ u64 ld_acq(const u64* p) { u64 v; asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ u64 ld_rlx(const u64* p) { u64 v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_rlx(u64* p, u64 v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ void st_rel(u64* p, u64 v) { asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ void red_rel(u64* p, u64 v) { asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ void red_rlx(u64* p, u64 v) { asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ u64 atom_add_rlx(u64* p, u64 v) { u64 o; asm volatile("atom.relaxed.gpu.global.add.u64 %0, [%1], %2;" : "=l"(o) : "l"(p), "l"(v) : "memory"); return o; }

struct Args { u64* w; float* bulk; int iters; int variant; int dirty; long long* out; };
// layout of w (u64 words, zero at launch): [0] flat counter; [64 + 16*g] group counters; [1024 + 16*r] release replicas;
// [4096 + q] LL slots word 0; [8192 + r*256 + q] replicated LL slots
__global__ void __launch_bounds__(256) k(Args a) {
  const int t = blockIdx.x, G = gridDim.x, tid = threadIdx.x;
  u64* w = a.w;
  long long t0 = 0;
  const long long t_start = clock64(), LIMIT = 4000000000ll;  // every spin gives up after ~2 s: a bug must not hang the GPU
  u64 target = 0;
  for (int it = 0; it < a.iters + 10; ++it) {
    if (it == 10) t0 = clock64();
    const u64 seq = (u64)it + 1;
    if (a.dirty) {  // every thread leaves stores in flight before the exchange (as P1 / P3 do)
      a.bulk[((size_t)t * 256 + tid) * 4 + (it & 3)] = (float)it;
      a.bulk[((size_t)((t * 7 + it) % G) * 256 + tid) * 4 + ((it + 1) & 3)] = (float)it;
    }
    __syncthreads();
    switch (a.variant) {
      case 0:  // flat counter: red.release + acquire polling by thread 0 (the library's grid_barrier)
        if (tid == 0) { target += G; red_rel(&w[0], 1); while (ld_acq(&w[0]) < target) { if (clock64() - t_start > LIMIT) break; } }
        break;
      case 1:  // flat counter, relaxed red after an explicit fence, relaxed polling + one acquire
        if (tid == 0) { target += G; __threadfence(); red_rlx(&w[0], 1); while (ld_rlx(&w[0]) < target) { if (clock64() - t_start > LIMIT) break; } (void)ld_acq(&w[0]); }
        break;
      case 2:  // flat counter, NO fence at all (lower bound of the counter scheme)
        if (tid == 0) { target += G; red_rlx(&w[0], 1); while (ld_rlx(&w[0]) < target) { if (clock64() - t_start > LIMIT) break; } }
        break;
      case 3: {  // two-level arrival (groups of 16), the last arriver releases through 16 replicated flags; no fence
        if (tid == 0) {
          const int g = t >> 4, ng = (G + 15) >> 4, gsz = min(16, G - g * 16);
          const u64 old = atom_add_rlx(&w[64 + 16 * g], 1);
          if ((old + 1) % gsz == 0) {
            const u64 o2 = atom_add_rlx(&w[32], 1);
            if ((o2 + 1) % ng == 0)
              for (int r = 0; r < 16; ++r) st_rlx(&w[1024 + 16 * r], seq);
          }
          while (ld_rlx(&w[1024 + 16 * (t & 15)]) < seq) { if (clock64() - t_start > LIMIT) break; }
        }
        break;
      }
      case 4: {  // as 3 with the release fence before arriving and an acquire at the end
        if (tid == 0) {
          const int g = t >> 4, ng = (G + 15) >> 4, gsz = min(16, G - g * 16);
          __threadfence();
          const u64 old = atom_add_rlx(&w[64 + 16 * g], 1);
          if ((old + 1) % gsz == 0) {
            const u64 o2 = atom_add_rlx(&w[32], 1);
            if ((o2 + 1) % ng == 0)
              for (int r = 0; r < 16; ++r) st_rlx(&w[1024 + 16 * r], seq);
          }
          while (ld_rlx(&w[1024 + 16 * (t & 15)]) < seq) { if (clock64() - t_start > LIMIT) break; }
          (void)ld_acq(&w[1024 + 16 * (t & 15)]);
        }
        break;
      }
      case 5: {  // flagged slots, all-to-all: every block stores one word, 32 threads poll 128 slots (4 each)
        if (tid == 0) st_rlx(&w[4096 + t], (seq << 32) | (u64)t);
        if (tid < 32)
          for (int q = tid; q < G; q += 32) while ((ld_rlx(&w[4096 + q]) >> 32) < seq) { if (clock64() - t_start > LIMIT) break; }
        break;
      }
      case 6: {  // flagged slots, 128 threads poll one slot each
        if (tid == 0) st_rlx(&w[4096 + t], (seq << 32) | (u64)t);
        for (int q = tid; q < G; q += 256) while ((ld_rlx(&w[4096 + q]) >> 32) < seq) { if (clock64() - t_start > LIMIT) break; }
        break;
      }
      case 7: {  // flagged slots replicated 16 times: block t reads replica t & 15 (8 readers per line instead of 128)
        if (tid < 16) st_rlx(&w[8192 + tid * 256 + t], (seq << 32) | (u64)t);
        if (tid >= 32 && tid < 64)
          for (int q = tid - 32; q < G; q += 32) while ((ld_rlx(&w[8192 + (t & 15) * 256 + q]) >> 32) < seq) { if (clock64() - t_start > LIMIT) break; }
        break;
      }
      case 8: {  // cooperative groups grid.sync()
        cg::this_grid().sync();
        break;
      }
      case 9: {  // flat counter, atom (returns) instead of red; the LAST arriver releases through 16 replicated flags; fence before
        if (tid == 0) {
          __threadfence();
          const u64 old = atom_add_rlx(&w[0], 1);
          if ((old + 1) % G == 0)
            for (int r = 0; r < 16; ++r) st_rlx(&w[1024 + 16 * r], seq);
          while (ld_rlx(&w[1024 + 16 * (t & 15)]) < seq) { if (clock64() - t_start > LIMIT) break; }
          (void)ld_acq(&w[1024 + 16 * (t & 15)]);
        }
        break;
      }
    }
    __syncthreads();
  }
  if (tid == 0) a.out[t] = clock64() - t0;
}

int main(int argc, char** argv) {
  const int G = argc > 1 ? atoi(argv[1]) : 128, iters = 2000;
  u64* w; float* bulk; long long* out;
  cudaMalloc(&w, 16384 * 8); cudaMalloc(&bulk, (size_t)G * 256 * 4 * 4); cudaMalloc(&out, G * 8);
  const char* names[] = {"flat counter: red.release + acquire polling (grid_barrier)", "flat counter: fence + relaxed red, relaxed polls + 1 acquire",
                         "flat counter: no fence (lower bound)", "two-level arrival + 16 replicated release flags, no fence",
                         "two-level arrival + replicated flags, fence + acquire", "flagged slots all-to-all, 32 pollers",
                         "flagged slots all-to-all, one slot per thread", "flagged slots replicated x16, 32 pollers", "cg grid.sync()",
                         "flat atom + last arriver releases 16 replicas, fence + acquire"};
  for (int dirty = 0; dirty < 2; ++dirty)
    for (int v = 0; v < 10; ++v) {
      cudaMemset(w, 0, 16384 * 8);
      Args a{w, bulk, iters, v, dirty, out};
      void* args[] = {&a};
      cudaError_t e = cudaLaunchCooperativeKernel((void*)k, dim3(G), dim3(256), args, 0, 0);
      if (e != cudaSuccess) { printf("launch: %s\n", cudaGetErrorString(e)); return 1; }
      e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("sync: %s\n", cudaGetErrorString(e)); return 1; }
      long long h[512];
      cudaMemcpy(h, out, G * 8, cudaMemcpyDeviceToHost);
      double mx = 0, mn = 1e30;
      for (int i = 0; i < G; ++i) { mx = h[i] > mx ? h[i] : mx; mn = h[i] < mn ? h[i] : mn; }
      printf("G=%d stores_in_flight=%d  %-70s %8.0f cycles per exchange (block min %.0f)\n", G, dirty, names[v], mx / iters, mn / iters);
    }
  return 0;
}
