// Micro-benchmark: dependent-issue latency of the instructions on the rANS decode step's critical path,
// one warp on one SM (clock64 around 512 dependent repetitions).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

#define REP 512
#define LAT_KERNEL(NAME, BODY)                                              \
  __global__ void NAME(uint32_t *out, uint32_t seed, long long *cyc) {       \
    __shared__ uint32_t tab[2048];                                          \
    for (int i = threadIdx.x; i < 2048; i += 32) tab[i] = (i * 2654435761u + seed) & 2047; \
    __syncwarp();                                                           \
    uint32_t v = (threadIdx.x * 37 + seed) & 2047, k = seed | 1, c = seed * 3 + 5;   \
    long long t0 = clock64();                                               \
    _Pragma("unroll 16") for (int i = 0; i < REP; ++i) { BODY }             \
    long long t1 = clock64();                                               \
    if (threadIdx.x == 0) *cyc = t1 - t0;                                   \
    out[threadIdx.x] = v + k + c + tab[0];                                  \
  }

LAT_KERNEL(l_lop3, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v) : "r"(k), "r"(c));)
LAT_KERNEL(l_imad, asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v) : "r"(k), "r"(c));)
LAT_KERNEL(l_imadhi, asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(v) : "r"(k)); v |= 0x80000000u;)
LAT_KERNEL(l_shf, asm volatile("shf.r.clamp.b32 %0, %0, %1, 3;" : "+r"(v) : "r"(k));)
LAT_KERNEL(l_popc, asm volatile("popc.b32 %0, %0;" : "+r"(v)); v += k;)
LAT_KERNEL(l_lds, v = tab[v & 2047];)
LAT_KERNEL(l_lds_u16, { uint32_t a = (uint32_t)__cvta_generic_to_shared(tab) + ((v & 2047) << 1); asm volatile("{ .reg .u16 t; ld.shared.u16 t, [%1]; cvt.u32.u16 %0, t; }" : "=r"(v) : "r"(a)); })
LAT_KERNEL(l_vote, { uint32_t m; asm volatile("{ .reg .pred p; setp.lt.u32 p, %1, %2; vote.sync.ballot.b32 %0, p, 0xffffffff; }" : "=r"(m) : "r"(v), "r"(c)); v += m; })
LAT_KERNEL(l_setp_sel, { uint32_t t; asm volatile("{ .reg .pred p; setp.lt.u32 p, %1, %2; selp.u32 %0, %1, %2, p; }" : "=r"(t) : "r"(v), "r"(c)); v = t + 1; })
LAT_KERNEL(l_vote_popc, { uint32_t m; asm volatile("{ .reg .pred p; setp.lt.u32 p, %1, %2; vote.sync.ballot.b32 %0, p, 0xffffffff; }" : "=r"(m) : "r"(v), "r"(c)); asm volatile("popc.b32 %0, %1;" : "=r"(m) : "r"(m)); v += m; })
LAT_KERNEL(l_shfl, asm volatile("shfl.sync.idx.b32 %0, %0, 3, 0x1f, 0xffffffff;" : "+r"(v)); v += k;)
LAT_KERNEL(l_leahi, asm volatile("{ .reg .u32 t; shr.u32 t, %0, 31; add.u32 %0, t, %1; }" : "+r"(v) : "r"(k));)
LAT_KERNEL(l_viaddmnmx, v = __viaddmin_s16x2_relu(v, k, c);)
// the decode step itself, one chain: table lookup -> update -> ballot -> popc -> word load -> renormalise
LAT_KERNEL(l_rans_step, {
  const uint32_t e = tab[v & 2047];
  uint32_t st = (v >> 11) * (e & 0xFFFu) + (e >> 20) + 0x7000u;
  const bool need = st < 0x8000u;
  const uint32_t m = __ballot_sync(0xffffffffu, need);
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(tab) + 4096u - 2u * __popc(m & k);
  uint32_t w; asm volatile("{ .reg .u16 t; ld.shared.u16 t, [%1]; cvt.u32.u16 %0, t; }" : "=r"(w) : "r"(a));
  if (need) st = (st << 16) | w;
  v = st;
})

// two independent chains in one instruction stream (what a decode warp runs), and the same with eight warps per CTA
#define RANS_STEP(v, off)                                                                         \
  {                                                                                               \
    const uint32_t e = tab[(v) & 2047];                                                           \
    uint32_t st = ((v) >> 11) * (e & 0xFFFu) + (e >> 20) + 0x7000u;                                \
    const bool need = st < 0x8000u;                                                               \
    const uint32_t m = __ballot_sync(0xffffffffu, need);                                          \
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(tab) + (off) - 2u * __popc(m & k);      \
    uint32_t w;                                                                                   \
    asm volatile("{ .reg .u16 t; ld.shared.u16 t, [%1]; cvt.u32.u16 %0, t; }" : "=r"(w) : "r"(a)); \
    if (need) st = (st << 16) | w;                                                                \
    (v) = st;                                                                                     \
  }
__global__ void l_rans_step2(uint32_t *out, uint32_t seed, long long *cyc) {
  __shared__ uint32_t tab[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) tab[i] = (i * 2654435761u + seed) & 2047;
  __syncthreads();
  uint32_t v = (threadIdx.x * 37 + seed) & 2047, v2 = (threadIdx.x * 41 + seed + 7) & 2047, k = seed | 1;
  long long t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < REP; ++i) {
    RANS_STEP(v, 4096u)
    RANS_STEP(v2, 6144u)
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) *cyc = t1 - t0;
  out[threadIdx.x] = v + v2 + k + tab[0];
}

typedef void (*kern_t)(uint32_t *, uint32_t, long long *);
int main() {
  uint32_t *out; long long *cyc, h;
  cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
  struct { const char *name; kern_t k; int ops; } ks[] = {
    {"LOP3", l_lop3, 1}, {"IMAD", l_imad, 1}, {"IMAD.HI (+LOP)", l_imadhi, 1}, {"SHF", l_shf, 1}, {"POPC (+IADD)", l_popc, 1},
    {"LDS.32 gather", l_lds, 1}, {"LDS.U16 (+addr)", l_lds_u16, 1}, {"SETP+VOTE (+IADD)", l_vote, 1}, {"SETP+SELP (+IADD)", l_setp_sel, 1},
    {"SETP+VOTE+POPC (+IADD)", l_vote_popc, 1}, {"SHFL (+IADD)", l_shfl, 1}, {"LEA.HI", l_leahi, 1}, {"VIADDMNMX", l_viaddmnmx, 1},
    {"rANS decode step (1 chain)", l_rans_step, 1}};
  printf("%-32s %s\n", "dependent chain", "cycles per repetition");
  for (auto &e : ks) {
    e.k<<<1, 32>>>(out, 1, cyc); cudaDeviceSynchronize();
    e.k<<<1, 32>>>(out, 1, cyc); cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-32s %8.1f\n", e.name, double(h) / REP);
  }
  for (int threads : {32, 128, 256, 512, 1024}) {
    l_rans_step2<<<1, threads>>>(out, 1, cyc); cudaDeviceSynchronize();
    l_rans_step2<<<1, threads>>>(out, 1, cyc); cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("rANS step, 2 chains, %4d threads   %8.1f\n", threads, double(h) / REP);
  }
  return 0;
}
