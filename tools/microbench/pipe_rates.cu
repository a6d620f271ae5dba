// Measures per-SM issue throughput (thread-ops / clk / SM) of the integer / DPX opcodes the
// Gotoh wavefront kernels are built from, on the GPU it runs on (sm_100a).  Not part of the
// product path: its output (profiles/r01_pipe_rates.txt) fixes the integer-issue roofline in DESIGN.md.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu && ./pipe_rates
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int ILP = 8;       // independent chains per thread
constexpr int UNROLL = 32;   // ops per chain per loop iteration

struct OpIadd3   { static constexpr const char* name = "IADD3 (x+y+z)";            __device__ static int f(int x, int y, int z) { return x + y + z; } };
struct OpImad    { static constexpr const char* name = "IMAD (x*y+z)";             __device__ static int f(int x, int y, int z) { return x * y + z; } };
struct OpLop3    { static constexpr const char* name = "LOP3 ((x&y)^z)";           __device__ static int f(int x, int y, int z) { return (x & y) ^ z; } };
struct OpPrmt    { static constexpr const char* name = "PRMT";                     __device__ static int f(int x, int y, int z) { return __byte_perm(x, y, z); } };
struct OpMax     { static constexpr const char* name = "VIMNMX s32 max";           __device__ static int f(int x, int y, int z) { return max(x, y); } };
struct OpMax16   { static constexpr const char* name = "VIMNMX.S16x2";             __device__ static int f(int x, int y, int z) { return __vmaxs2(x, y); } };
struct OpMax3    { static constexpr const char* name = "VIMNMX3 s32";              __device__ static int f(int x, int y, int z) { return __vimax3_s32(x, y, z); } };
struct OpMax3_16 { static constexpr const char* name = "VIMNMX3.S16x2";            __device__ static int f(int x, int y, int z) { return __vimax3_s16x2(x, y, z); } };
struct OpAddMax  { static constexpr const char* name = "VIADDMNMX s32";            __device__ static int f(int x, int y, int z) { return __viaddmax_s32(x, y, z); } };
struct OpAddMaxI { static constexpr const char* name = "VIADDMNMX s32 imm";        __device__ static int f(int x, int y, int z) { return __viaddmax_s32(x, -2, z); } };
struct OpAddMaxR { static constexpr const char* name = "VIADDMNMX.RELU s32";       __device__ static int f(int x, int y, int z) { return __viaddmax_s32_relu(x, y, z); } };
struct OpAddMax16{ static constexpr const char* name = "VIADDMNMX.S16x2";          __device__ static int f(int x, int y, int z) { return __viaddmax_s16x2(x, y, z); } };
struct OpAddMax16I{static constexpr const char* name = "VIADDMNMX.S16x2 imm";      __device__ static int f(int x, int y, int z) { return __viaddmax_s16x2(x, 0xFFFEFFFEu, z); } };
struct OpAddMax16R{static constexpr const char* name = "VIADDMNMX.S16x2.RELU";     __device__ static int f(int x, int y, int z) { return __viaddmax_s16x2_relu(x, y, z); } };
struct OpVadd2   { static constexpr const char* name = "VIADD.16x2";               __device__ static int f(int x, int y, int z) { return __vadd2(x, y); } };
struct OpVadd2I  { static constexpr const char* name = "VIADD.16x2 imm";           __device__ static int f(int x, int y, int z) { return __vadd2(x, 0xFFFBFFFBu); } };
struct OpShfl    { static constexpr const char* name = "SHFL.UP";                  __device__ static int f(int x, int y, int z) { return __shfl_up_sync(0xffffffffu, x, 1); } };
struct OpSetSel  { static constexpr const char* name = "ISETP+SEL (2 instr)";      __device__ static int f(int x, int y, int z) { return (x == y) ? z : x + 1; } };
struct OpFfma    { static constexpr const char* name = "FFMA";                     __device__ static int f(int x, int y, int z) { return __float_as_int(fmaf(__int_as_float(x), __int_as_float(y), __int_as_float(z))); } };
// pairs: alternate two opcodes to see whether they issue down different pipes
struct OpMixAddMaxImad   { static constexpr const char* name = "mix VIADDMNMX.S16x2 + IMAD (2 instr)";   __device__ static int f(int x, int y, int z) { return __viaddmax_s16x2(x * y + z, y, z); } };
struct OpMixMaxImad      { static constexpr const char* name = "mix VIMNMX + IMAD (2 instr)";            __device__ static int f(int x, int y, int z) { return max(x * y + z, z); } };
struct OpMixMaxIadd      { static constexpr const char* name = "mix VIMNMX + IADD3 (2 instr)";           __device__ static int f(int x, int y, int z) { return max(x + y + z, z); } };
struct OpMixPrmtAddMax   { static constexpr const char* name = "mix PRMT + VIADDMNMX.S16x2 (2 instr)";   __device__ static int f(int x, int y, int z) { return __viaddmax_s16x2(__byte_perm(x, y, z), y, z); } };
struct OpMixAddMaxFfma   { static constexpr const char* name = "mix VIADDMNMX.S16x2 + FFMA (2 instr)";   __device__ static int f(int x, int y, int z) { return __viaddmax_s16x2(__float_as_int(fmaf(__int_as_float(x), 1.0001f, 0.5f)), y, z); } };
struct OpMixMax3Vadd     { static constexpr const char* name = "mix VIMNMX3.S16x2 + VIADD.16x2 (2 instr)";__device__ static int f(int x, int y, int z) { return __vadd2(__vimax3_s16x2(x, y, z), 0xFFFBFFFBu); } };

template <class Op>
__global__ void __launch_bounds__(256) bench(int* out, int iters, int y0, int z0, long long* cyc) {
    int x[ILP];
    int y = y0 + (threadIdx.x & 3); (void)y;
#pragma unroll
    for (int k = 0; k < ILP; k++) x[k] = threadIdx.x * 7 + k;
    long long c0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            int n[ILP];
#pragma unroll
            for (int k = 0; k < ILP; k++) n[k] = Op::f(x[k], x[(k + 3) % ILP], x[(k + 5) % ILP]);
#pragma unroll
            for (int k = 0; k < ILP; k++) x[k] = n[k];
        }
    }
    long long c1 = clock64();
    int acc = 0;
#pragma unroll
    for (int k = 0; k < ILP; k++) acc ^= x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
}

template <class Op>
void run(int sms, int blocks_per_sm, int* d_out, long long* d_cyc, int instr_per_op) {
    int iters = 2000;
    int grid = sms * blocks_per_sm;
    bench<Op><<<grid, 256>>>(d_out, 10, 3, 1000, d_cyc);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    bench<Op><<<grid, 256>>>(d_out, iters, 3, 1000, d_cyc);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    long long* h = (long long*)malloc(grid * sizeof(long long));
    CK(cudaMemcpy(h, d_cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
    double cmax = 0; for (int i = 0; i < grid; i++) if (h[i] > cmax) cmax = (double)h[i];
    free(h);
    double ops = (double)iters * UNROLL * ILP * 256.0 * blocks_per_sm;   // thread-ops per SM
    printf("%-44s bps=%d  %8.2f thread-ops/clk/SM  (%6.2f thread-instr/clk/SM)  %7.3f ms  %.0f MHz eff\n", Op::name, blocks_per_sm,
           ops / cmax, ops * instr_per_op / cmax, ms, cmax / (ms * 1e3));
    CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("# %s  SMs=%d  cc=%d.%d  clock=%d kHz\n", p.name, sms, p.major, p.minor, p.clockRate);
    int* d_out; long long* d_cyc;
    CK(cudaMalloc(&d_out, sizeof(int) * sms * 8 * 256)); CK(cudaMalloc(&d_cyc, sizeof(long long) * sms * 8));
    for (int bps : {2, 4}) {   // 16 and 32 warps per SM
        run<OpIadd3>(sms, bps, d_out, d_cyc, 1);
        run<OpImad>(sms, bps, d_out, d_cyc, 1);
        run<OpFfma>(sms, bps, d_out, d_cyc, 1);
        run<OpLop3>(sms, bps, d_out, d_cyc, 1);
        run<OpPrmt>(sms, bps, d_out, d_cyc, 1);
        run<OpMax>(sms, bps, d_out, d_cyc, 1);
        run<OpMax16>(sms, bps, d_out, d_cyc, 1);
        run<OpMax3>(sms, bps, d_out, d_cyc, 1);
        run<OpMax3_16>(sms, bps, d_out, d_cyc, 1);
        run<OpAddMax>(sms, bps, d_out, d_cyc, 1);
        run<OpAddMaxI>(sms, bps, d_out, d_cyc, 1);
        run<OpAddMaxR>(sms, bps, d_out, d_cyc, 1);
        run<OpAddMax16>(sms, bps, d_out, d_cyc, 1);
        run<OpAddMax16I>(sms, bps, d_out, d_cyc, 1);
        run<OpAddMax16R>(sms, bps, d_out, d_cyc, 1);
        run<OpVadd2>(sms, bps, d_out, d_cyc, 1);
        run<OpVadd2I>(sms, bps, d_out, d_cyc, 1);
        run<OpSetSel>(sms, bps, d_out, d_cyc, 2);
        run<OpShfl>(sms, bps, d_out, d_cyc, 1);
        run<OpMixAddMaxImad>(sms, bps, d_out, d_cyc, 2);
        run<OpMixMaxImad>(sms, bps, d_out, d_cyc, 2);
        run<OpMixMaxIadd>(sms, bps, d_out, d_cyc, 2);
        run<OpMixPrmtAddMax>(sms, bps, d_out, d_cyc, 2);
        run<OpMixAddMaxFfma>(sms, bps, d_out, d_cyc, 2);
        run<OpMixMax3Vadd>(sms, bps, d_out, d_cyc, 2);
    }
    return 0;
}
