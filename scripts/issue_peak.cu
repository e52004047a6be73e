// issue_peak.cu -- measurement aid: the integer issue ceiling and the dependent-op latencies
// that bound the coder kernels, measured on the box.  Standalone: built by scripts/build_issue_peak.sh, run
// under gpurun, prints one JSON object.  Not part of libredux_b200.so.
//
//   issue peak    all SMs, 32 warps per SM, per thread 8 independent chains of the integer mix the coders
//                 use (IADD3 / LOP3 / SHF on the ALU pipe, IMAD on the FMA pipe): warp-instructions per second
//   latencies     one warp, one long dependent chain per op: cycles from issue to a usable result
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(2); } } while (0)

constexpr int kIters = 4096;

// mix: 0 = ALU only (IADD3, LOP3, SHF), 1 = IMAD only, 2 = alternating (half ALU, half IMAD)
template <int MIX>
__global__ void __launch_bounds__(1024) issue_kernel(unsigned *out, unsigned seed)
{
    unsigned a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 8 + i;
    const unsigned m = seed | 1u;
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MIX == 0) {
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(m));
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(a[i]) : "r"(m));
                asm volatile("shf.l.wrap.b32 %0, %0, %0, 3;" : "+r"(a[i]));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(m));
            } else if (MIX == 1) {
                asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(a[i]) : "r"(m));
                asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(a[i]) : "r"(m));
                asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(a[i]) : "r"(m));
                asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(a[i]) : "r"(m));
            } else {
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(m));
                asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(a[i]) : "r"(m));
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(a[i]) : "r"(m));
                asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(a[i]) : "r"(m));
            }
        }
    }
    unsigned r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r ^= a[i];
    if (r == 0x12345678u) out[0] = r;           // keep the chains alive
}

enum Op { kAdd, kLop, kShf, kMad, kMadHi, kFlo, kLds, kShfl, kRedux, kPrmt, kSel, kNumOps };
static const char *kOpNames[kNumOps] = {"IADD3/LOP3 alternating", "LOP3/IADD3 alternating", "SHF", "IMAD", "IMAD.HI", "FLO.SH", "LDS", "SHFL", "REDUX.MAX",
                                        "PRMT", "ISETP+SEL"};

template <int OP>
__global__ void latency_kernel(long long *cycles, unsigned *out, unsigned seed)
{
    __shared__ unsigned chase[64];
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(chase);
    for (int i = threadIdx.x; i < 64; i += 32) chase[i] = sbase + ((i + 7) & 63) * 4;   // shared addresses of a 64-cycle
    __syncwarp();
    unsigned x = OP == kLds ? sbase + threadIdx.x * 4 : seed + threadIdx.x, m = seed | 1u;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kIters / 16; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            // (a chain of identical adds or xors is fused by ptxas into 3-input IADD3 / LOP3: alternate them)
            if (OP == kAdd) { if (u & 1) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(m));
                              else asm volatile("xor.b32 %0, %0, %1;" : "+r"(x) : "r"(m)); }
            if (OP == kLop) { if (u & 1) asm volatile("and.b32 %0, %0, %1;" : "+r"(x) : "r"(~m));
                              else asm volatile("sub.u32 %0, %1, %0;" : "+r"(x) : "r"(m)); }
            if (OP == kShf) asm volatile("shf.l.wrap.b32 %0, %0, %0, 3;" : "+r"(x));
            if (OP == kMad) asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(x) : "r"(m));
            if (OP == kMadHi) asm volatile("mad.hi.u32 %0, %0, %1, %1;" : "+r"(x) : "r"(m));
            if (OP == kFlo) { unsigned y; asm volatile("bfind.shiftamt.u32 %0, %1;" : "=r"(y) : "r"(x));
                              asm volatile("add.u32 %0, %1, %2;" : "=r"(x) : "r"(y), "r"(m)); }        // + one IADD
            if (OP == kLds) asm volatile("ld.shared.u32 %0, [%0];" : "+r"(x));
            if (OP == kShfl) x = __shfl_sync(0xFFFFFFFFu, x, (threadIdx.x + 1) & 31);
            if (OP == kRedux) x = __reduce_max_sync(0xFFFFFFFFu, x) + threadIdx.x;            // + one IADD
            if (OP == kPrmt) asm volatile("prmt.b32 %0, %0, %1, 0x0123;" : "+r"(x) : "r"(m));
            if (OP == kSel) x = (x > m) ? x - m : x + 1;                                        // ISETP + 2 ALU
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
    if (x == 0x12345678u) out[0] = x;
}

template <int OP>
double latency(long long *d_cycles, unsigned *d_out, bool lds)
{
    long long h = 0;
    latency_kernel<OP><<<1, 32>>>(d_cycles, d_out, lds ? 0u : 12345u);
    CHECK(cudaDeviceSynchronize());
    latency_kernel<OP><<<1, 32>>>(d_cycles, d_out, lds ? 0u : 12345u);
    CHECK(cudaDeviceSynchronize());
    CHECK(cudaMemcpy(&h, d_cycles, sizeof h, cudaMemcpyDeviceToHost));
    return (double)h / kIters;
}

template <int MIX>
double issue_rate(unsigned *d_out, int sms)
{
    cudaEvent_t a, b;
    CHECK(cudaEventCreate(&a)); CHECK(cudaEventCreate(&b));
    const int grid = sms * 2;                              // 2 x 1024 threads = 64 warps per SM
    issue_kernel<MIX><<<grid, 1024>>>(d_out, 12345u);      // warm-up
    CHECK(cudaDeviceSynchronize());
    CHECK(cudaEventRecord(a));
    issue_kernel<MIX><<<grid, 1024>>>(d_out, 12345u);
    CHECK(cudaEventRecord(b));
    CHECK(cudaEventSynchronize(b));
    float ms = 0;
    CHECK(cudaEventElapsedTime(&ms, a, b));
    const double inst = (double)grid * 32 /*warps*/ * kIters * 8 * 4;
    return inst / (ms * 1e-3);
}

int main()
{
    cudaDeviceProp p;
    CHECK(cudaGetDeviceProperties(&p, 0));
    unsigned *d_out; long long *d_cycles;
    CHECK(cudaMalloc(&d_out, 64)); CHECK(cudaMalloc(&d_cycles, 64));
    const double alu = issue_rate<0>(d_out, p.multiProcessorCount);
    const double mad = issue_rate<1>(d_out, p.multiProcessorCount);
    const double mix = issue_rate<2>(d_out, p.multiProcessorCount);
    int clock_khz = 0;
    CHECK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0));
    const double nominal = (double)p.multiProcessorCount * 4 * clock_khz * 1e3;
    double lat[kNumOps];
    lat[kAdd] = latency<kAdd>(d_cycles, d_out, false);   lat[kLop] = latency<kLop>(d_cycles, d_out, false);
    lat[kShf] = latency<kShf>(d_cycles, d_out, false);   lat[kMad] = latency<kMad>(d_cycles, d_out, false);
    lat[kMadHi] = latency<kMadHi>(d_cycles, d_out, false); lat[kFlo] = latency<kFlo>(d_cycles, d_out, false);
    lat[kLds] = latency<kLds>(d_cycles, d_out, true);    lat[kShfl] = latency<kShfl>(d_cycles, d_out, false);
    lat[kRedux] = latency<kRedux>(d_cycles, d_out, false); lat[kPrmt] = latency<kPrmt>(d_cycles, d_out, false);
    lat[kSel] = latency<kSel>(d_cycles, d_out, false);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_clock_mhz\": %.0f,\n", p.name, p.multiProcessorCount, clock_khz / 1e3);
    printf(" \"issue_nominal_warp_inst_per_s\": %.4g, \"note_nominal\": \"SMs x 4 schedulers x 1 warp-inst/clk x max SM clock\",\n", nominal);
    printf(" \"issue_measured_warp_inst_per_s\": {\"alu_only\": %.4g, \"imad_only\": %.4g, \"alu_imad_mix\": %.4g},\n", alu, mad, mix);
    printf(" \"issue_measured_per_sm_per_clk\": {\"alu_only\": %.3f, \"imad_only\": %.3f, \"alu_imad_mix\": %.3f},\n",
           alu / (p.multiProcessorCount * clock_khz * 1e3), mad / (p.multiProcessorCount * clock_khz * 1e3),
           mix / (p.multiProcessorCount * clock_khz * 1e3));
    printf(" \"dependent_latency_cycles\": {");
    for (int i = 0; i < kNumOps; ++i) printf("%s\"%s\": %.2f", i ? ", " : "", kOpNames[i], lat[i]);
    printf("},\n \"latency_notes\": \"one warp, one dependent chain; FLO.SH includes one IADD3, REDUX.MAX one IADD3, ISETP+SEL is compare + two ALU ops\"}\n");
    return 0;
}
