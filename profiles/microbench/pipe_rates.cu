// Micro-benchmark: sustained throughput of the integer instructions the decode kernels lean on,
// per SM per clock on B200 (sm_100a).  Each kernel runs 8 independent dependency chains per
// thread, 1024 threads per block, 2 blocks per SM, so the issue ports -- not latency -- bound it.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

#define ITERS 4096
#define ILP 8

#define BENCH_KERNEL(NAME, BODY)                                                    \
  __global__ void __launch_bounds__(1024) NAME(uint32_t *out, uint32_t seed) {      \
    uint32_t v[ILP];                                                                \
    uint32_t k = seed + threadIdx.x, c = seed * 3 + 1;                              \
    _Pragma("unroll") for (int j = 0; j < ILP; ++j) v[j] = k * (j + 1) + seed;      \
    for (int i = 0; i < ITERS; ++i) {                                               \
      _Pragma("unroll") for (int j = 0; j < ILP; ++j) { BODY }                      \
    }                                                                               \
    uint32_t s = 0;                                                                 \
    _Pragma("unroll") for (int j = 0; j < ILP; ++j) s ^= v[j];                      \
    if (s == 0x12345678u) out[threadIdx.x] = s + k + c;                             \
  }

BENCH_KERNEL(k_lop3, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[j]) : "r"(k), "r"(c));)
BENCH_KERNEL(k_shf, asm volatile("shr.u32 %0, %0, 1;" : "+r"(v[j])); asm volatile("shl.b32 %0, %0, 1;" : "+r"(v[j]));)
BENCH_KERNEL(k_shr, asm volatile("shf.r.clamp.b32 %0, %0, %1, 3;" : "+r"(v[j]) : "r"(k));)
BENCH_KERNEL(k_iadd, asm volatile("add.u32 %0, %0, %1;" : "+r"(v[j]) : "r"(k));)
BENCH_KERNEL(k_imad, asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[j]) : "r"(k), "r"(c));)
BENCH_KERNEL(k_imadhi, asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(v[j]) : "r"(k));)
BENCH_KERNEL(k_prmt, asm volatile("prmt.b32 %0, %0, %1, 0x2104;" : "+r"(v[j]) : "r"(k));)
BENCH_KERNEL(k_popc, asm volatile("popc.b32 %0, %0;" : "+r"(v[j])); v[j] += k;)
BENCH_KERNEL(k_setp, { uint32_t t; asm volatile("{ .reg .pred p; setp.lt.u32 p, %1, %2; selp.u32 %0, %1, %2, p; }" : "=r"(t) : "r"(v[j]), "r"(c)); v[j] = t + 1; })
BENCH_KERNEL(k_vote, { uint32_t t; asm volatile("{ .reg .pred p; setp.lt.u32 p, %1, %2; vote.sync.ballot.b32 %0, p, 0xffffffff; }" : "=r"(t) : "r"(v[j]), "r"(c)); v[j] += t; })
BENCH_KERNEL(k_shfl, asm volatile("shfl.sync.idx.b32 %0, %0, 0, 0x1f, 0xffffffff;" : "+r"(v[j])); v[j] += k;)
BENCH_KERNEL(k_mix_lop_imad, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[j]) : "r"(k), "r"(c)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[j]) : "r"(k), "r"(c));)
BENCH_KERNEL(k_mix_lop_imadhi, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[j]) : "r"(k), "r"(c)); asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(v[j]) : "r"(k));)
BENCH_KERNEL(k_mix_lop_shf, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[j]) : "r"(k), "r"(c)); asm volatile("shf.r.clamp.b32 %0, %0, %1, 3;" : "+r"(v[j]) : "r"(k));)
BENCH_KERNEL(k_mix_lop_popc, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[j]) : "r"(k), "r"(c)); asm volatile("popc.b32 %0, %0;" : "+r"(v[j]));)
BENCH_KERNEL(k_ffma, { float f = __uint_as_float(v[j]); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(1.0001f), "f"(0.5f)); v[j] = __float_as_uint(f); })
BENCH_KERNEL(k_mix_lop_ffma, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[j]) : "r"(k), "r"(c)); { float f = __uint_as_float(v[j]); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(1.0001f), "f"(0.5f)); v[j] = __float_as_uint(f); })
BENCH_KERNEL(k_hadd2, asm volatile("add.f16x2 %0, %0, %1;" : "+r"(v[j]) : "r"(k));)
BENCH_KERNEL(k_dp4a, asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(v[j]) : "r"(k), "r"(c));)
BENCH_KERNEL(k_vabsdiff, asm volatile("vadd2.s32.s32.s32 %0, %0, %1, %2;" : "+r"(v[j]) : "r"(k), "r"(c));)

BENCH_KERNEL(k_dp2a, asm volatile("dp2a.lo.s32.s32 %0, %0, %1, %2;" : "+r"(v[j]) : "r"(k), "r"(c));)
BENCH_KERNEL(k_mix_lop_dp4a, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[j]) : "r"(k), "r"(c)); asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(v[j]) : "r"(k), "r"(c));)
BENCH_KERNEL(k_mix_imad_dp4a, asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[j]) : "r"(k), "r"(c)); asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(v[j]) : "r"(k), "r"(c));)
BENCH_KERNEL(k_mix_lop_iadd3, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[j]) : "r"(k), "r"(c)); asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(v[j]) : "r"(k), "r"(c));)
BENCH_KERNEL(k_mix_lop_hadd2, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[j]) : "r"(k), "r"(c)); asm volatile("add.f16x2 %0, %0, %1;" : "+r"(v[j]) : "r"(k));)
BENCH_KERNEL(k_mix_imad_hadd2, asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[j]) : "r"(k), "r"(c)); asm volatile("add.f16x2 %0, %0, %1;" : "+r"(v[j]) : "r"(k));)
BENCH_KERNEL(k_mix_lop_shfl, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[j]) : "r"(k), "r"(c)); asm volatile("shfl.sync.idx.b32 %0, %0, 0, 0x1f, 0xffffffff;" : "+r"(v[j]));)
BENCH_KERNEL(k_mix3, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[j]) : "r"(k), "r"(c)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[j]) : "r"(k), "r"(c)); { float f = __uint_as_float(v[j]); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(1.0001f), "f"(0.5f)); v[j] = __float_as_uint(f); })
BENCH_KERNEL(k_lea, asm volatile("{ .reg .u32 t; shl.b32 t, %0, 3; add.u32 %0, t, %1; }" : "+r"(v[j]) : "r"(k));)
BENCH_KERNEL(k_viadd16x2, asm volatile("add.s16x2 %0, %0, %1;" : "+r"(v[j]) : "r"(k));)
BENCH_KERNEL(k_vimnmx16x2, asm volatile("max.s16x2 %0, %0, %1;" : "+r"(v[j]) : "r"(k));)
BENCH_KERNEL(k_mix_lop_viadd16x2, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[j]) : "r"(k), "r"(c)); asm volatile("add.s16x2 %0, %0, %1;" : "+r"(v[j]) : "r"(k));)
BENCH_KERNEL(k_mix_imad_viadd16x2, asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[j]) : "r"(k), "r"(c)); asm volatile("add.s16x2 %0, %0, %1;" : "+r"(v[j]) : "r"(k));)
BENCH_KERNEL(k_mix_lop_vimnmx16x2, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[j]) : "r"(k), "r"(c)); asm volatile("max.s16x2 %0, %0, %1;" : "+r"(v[j]) : "r"(k));)
BENCH_KERNEL(k_iadd3, asm volatile("{ .reg .u32 t; add.u32 t, %1, %2; add.u32 %0, %0, t; }" : "+r"(v[j]) : "r"(k), "r"(c)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[j]) : "r"(k), "r"(c));)
BENCH_KERNEL(k_imax, asm volatile("max.s32 %0, %0, %1;" : "+r"(v[j]) : "r"(k));)
BENCH_KERNEL(k_sar_fix, { int x = (int)v[j]; x = x / 4; v[j] = (uint32_t)x + k; })

BENCH_KERNEL(k_viaddmnmx, v[j] = __viaddmin_s16x2_relu(v[j], k, c);)
BENCH_KERNEL(k_viaddmnmx_u, v[j] = __viaddmin_u16x2(v[j], k, c);)
BENCH_KERNEL(k_vimnmx3, v[j] = __vimax3_s16x2(v[j], k, c);)
BENCH_KERNEL(k_mix_lop_viaddmnmx, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[j]) : "r"(k), "r"(c)); v[j] = __viaddmin_s16x2_relu(v[j], k, c);)
BENCH_KERNEL(k_mix_imad_viaddmnmx, asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[j]) : "r"(k), "r"(c)); v[j] = __viaddmin_s16x2_relu(v[j], k, c);)
BENCH_KERNEL(k_mix_iadd_viaddmnmx, asm volatile("add.u32 %0, %0, %1;" : "+r"(v[j]) : "r"(k)); v[j] = __viaddmin_s16x2_relu(v[j], k, c);)
BENCH_KERNEL(k_mix_lop_vimnmx3, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[j]) : "r"(k), "r"(c)); v[j] = __vimax3_s16x2(v[j], k, c);)
BENCH_KERNEL(k_iadd3_only, asm volatile("{ .reg .u32 t; add.u32 t, %1, %2; add.u32 %0, %0, t; }" : "+r"(v[j]) : "r"(k), "r"(c));)
BENCH_KERNEL(k_lea_hi, asm volatile("{ .reg .u32 t; shr.u32 t, %0, 31; add.u32 %0, t, %1; }" : "+r"(v[j]) : "r"(k));)
BENCH_KERNEL(k_mix_lop_lea, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[j]) : "r"(k), "r"(c)); asm volatile("{ .reg .u32 t; shl.b32 t, %0, 3; add.u32 %0, t, %1; }" : "+r"(v[j]) : "r"(k));)
BENCH_KERNEL(k_mix_imad_iadd, asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[j]) : "r"(k), "r"(c)); asm volatile("add.u32 %0, %0, %1;" : "+r"(v[j]) : "r"(k));)
BENCH_KERNEL(k_mix_lop_iadd_imad, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[j]) : "r"(k), "r"(c)); asm volatile("add.u32 %0, %0, %1;" : "+r"(v[j]) : "r"(k)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[j]) : "r"(k), "r"(c));)
BENCH_KERNEL(k_isetp_only, { uint32_t t; asm volatile("{ .reg .pred p; setp.lt.u32 p, %1, %2; selp.u32 %0, %1, %2, p; }" : "=r"(t) : "r"(v[j]), "r"(c)); v[j] = t; })

BENCH_KERNEL(k_imadhi_3iadd, asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(v[j]) : "r"(k)); asm volatile("add.u32 %0, %0, %1;" : "+r"(v[j]) : "r"(k)); asm volatile("add.u32 %0, %0, %1;" : "+r"(v[j]) : "r"(c)); asm volatile("add.u32 %0, %0, %1;" : "+r"(v[j]) : "r"(k));)
BENCH_KERNEL(k_imadhi_2lop, asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(v[j]) : "r"(k)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[j]) : "r"(k), "r"(c)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[j]) : "r"(c), "r"(k));)
BENCH_KERNEL(k_imadhi_2imad, asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(v[j]) : "r"(k)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[j]) : "r"(k), "r"(c)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[j]) : "r"(c), "r"(k));)
BENCH_KERNEL(k_imadwide, { unsigned long long w; asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(w) : "r"(v[j]), "r"(k), "l"((unsigned long long)c)); v[j] = (uint32_t)w ^ (uint32_t)(w >> 32); })
BENCH_KERNEL(k_imadwide_2lop, { unsigned long long w; asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(w) : "r"(v[j]), "r"(k), "l"((unsigned long long)c)); v[j] = (uint32_t)w + (uint32_t)(w >> 32); } asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[j]) : "r"(k), "r"(c));)
BENCH_KERNEL(k_leahi_imad, asm volatile("{ .reg .u32 t; shr.u32 t, %0, 31; add.u32 %0, t, %1; }" : "+r"(v[j]) : "r"(k)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[j]) : "r"(k), "r"(c));)

__global__ void __launch_bounds__(1024) k_lds(uint32_t *out, uint32_t seed) {
  __shared__ uint32_t tab[2048];
  for (int i = threadIdx.x; i < 2048; i += 1024) tab[i] = (i * 2654435761u + seed) & 2047;
  __syncthreads();
  uint32_t v[ILP];
#pragma unroll
  for (int j = 0; j < ILP; ++j) v[j] = (threadIdx.x * 37 + j * 101 + seed) & 2047;
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int j = 0; j < ILP; ++j) v[j] = tab[v[j]];
  }
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < ILP; ++j) s ^= v[j];
  if (s == 0x12345678u) out[threadIdx.x] = s;
}

typedef void (*kern_t)(uint32_t *, uint32_t);

int main() {
  uint32_t *out;
  cudaMalloc(&out, 4096 * 4);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  struct { const char *name; kern_t k; int ops; } ks[] = {
      {"LOP3", k_lop3, 1}, {"SHR+SHL (2 ops)", k_shf, 2}, {"SHF.R funnel", k_shr, 1}, {"IADD", k_iadd, 1},
      {"IMAD.lo", k_imad, 1}, {"IMAD.HI (mul.hi)", k_imadhi, 1}, {"PRMT", k_prmt, 1}, {"POPC (+IADD)", k_popc, 2},
      {"SETP+SELP (+IADD)", k_setp, 3}, {"SETP+VOTE.ballot (+IADD)", k_vote, 3}, {"SHFL.idx (+IADD)", k_shfl, 2},
      {"LOP3+IMAD alternating", k_mix_lop_imad, 2}, {"LOP3+IMAD.HI alternating", k_mix_lop_imadhi, 2},
      {"LOP3+SHF alternating", k_mix_lop_shf, 2}, {"LOP3+POPC alternating", k_mix_lop_popc, 2}, {"FFMA", k_ffma, 1},
      {"LOP3+FFMA alternating", k_mix_lop_ffma, 2}, {"HADD2 (add.f16x2)", k_hadd2, 1}, {"DP4A", k_dp4a, 1},
      {"vadd2 (SIMD video, emulated?)", k_vabsdiff, 1}, {"LDS random 2048-entry gather", k_lds, 1},
      {"DP2A", k_dp2a, 1}, {"LOP3+DP4A alternating", k_mix_lop_dp4a, 2}, {"IMAD+DP4A alternating", k_mix_imad_dp4a, 2},
      {"LOP3+2xIADD", k_mix_lop_iadd3, 3}, {"LOP3+HADD2 alternating", k_mix_lop_hadd2, 2},
      {"IMAD+HADD2 alternating", k_mix_imad_hadd2, 2}, {"LOP3+SHFL alternating", k_mix_lop_shfl, 2},
      {"LOP3+IMAD+FFMA", k_mix3, 3}, {"VIADD.16x2 (add.s16x2)", k_viadd16x2, 1}, {"VIMNMX.16x2 (max.s16x2)", k_vimnmx16x2, 1},
      {"LOP3+VIADD.16x2 alternating", k_mix_lop_viadd16x2, 2}, {"IMAD+VIADD.16x2 alternating", k_mix_imad_viadd16x2, 2},
      {"LOP3+VIMNMX.16x2 alternating", k_mix_lop_vimnmx16x2, 2}, {"VIMNMX (max.s32)", k_imax, 1}, {"SHL+ADD (LEA?)", k_lea, 1}, {"x/4 signed (+IADD)", k_sar_fix, 1},
      {"VIADDMNMX.S16x2.RELU", k_viaddmnmx, 1}, {"VIADDMNMX.U16x2", k_viaddmnmx_u, 1}, {"VIMNMX3.S16x2", k_vimnmx3, 1},
      {"LOP3+VIADDMNMX alternating", k_mix_lop_viaddmnmx, 2}, {"IMAD+VIADDMNMX alternating", k_mix_imad_viaddmnmx, 2},
      {"IADD+VIADDMNMX alternating", k_mix_iadd_viaddmnmx, 2}, {"LOP3+VIMNMX3 alternating", k_mix_lop_vimnmx3, 2},
      {"IADD3 (2 adds)", k_iadd3_only, 1}, {"SHR31+ADD (LEA.HI?)", k_lea_hi, 1}, {"LOP3+LEA alternating", k_mix_lop_lea, 2},
      {"IMAD+IADD alternating", k_mix_imad_iadd, 2}, {"LOP3+IADD+IMAD", k_mix_lop_iadd_imad, 3}, {"SETP+SELP", k_isetp_only, 2},
      {"IMAD.HI+3xIADD", k_imadhi_3iadd, 4}, {"IMAD.HI+2xLOP3", k_imadhi_2lop, 3}, {"IMAD.HI+2xIMAD", k_imadhi_2imad, 3},
      {"IMAD.WIDE (+LOP3)", k_imadwide, 2}, {"IMAD.WIDE+IADD+LOP3", k_imadwide_2lop, 3}, {"LEA.HI+IMAD alternating", k_leahi_imad, 2}};
  printf("device %s, %d SMs, clock attr %d kHz\n", prop.name, sms, khz);
  printf("%-34s %12s %16s\n", "instruction", "ms", "thread-ops/clk/SM");
  for (auto &e : ks) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    e.k<<<sms * 2, 1024>>>(out, 1);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    e.k<<<sms * 2, 1024>>>(out, 1);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    double ops = double(sms) * 2 * 1024 * ITERS * ILP * e.ops;
    double per_clk_sm = ops / (ms * 1e-3) / (double(khz) * 1e3) / sms;
    printf("%-34s %12.4f %16.1f\n", e.name, ms, per_clk_sm);
  }
  return 0;
}
