// Per-instruction issue cost on one SM sub-partition (bring-up tool): cycles per warp instruction for the
// instructions the softmax warps are made of, at 1 / 2 / 4 warps per scheduler, alone and in pairs (to see which
// of them share a pipe). Eight independent dependency chains per thread, so latency is hidden from 1 warp up.
//   nvcc -O3 -arch=sm_100a -o ubench_pipes ubench_pipes.cu && ./ubench_pipes
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define DEV __device__ __forceinline__
DEV void op_ffma(float& a, float b, float c) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a) : "f"(b), "f"(c)); }
DEV void op_ffma_imm(float& a, float c) { asm volatile("fma.rn.f32 %0, %0, 0f3F7FF000, %1;" : "+f"(a) : "f"(c)); }
DEV void op_fadd(float& a, float b) { asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a) : "f"(b)); }
DEV void op_ffma2(float& a0, float& a1, float b, float c) {
    asm volatile("{\n\t.reg .b64 x, y, z;\n\tmov.b64 x, {%0, %1};\n\tmov.b64 y, {%2, %2};\n\tmov.b64 z, {%3, %3};\n\t"
                 "fma.rn.f32x2 x, x, y, z;\n\tmov.b64 {%0, %1}, x;\n\t}" : "+f"(a0), "+f"(a1) : "f"(b), "f"(c));
}
DEV void op_fadd2(float& a0, float& a1, float b) {
    asm volatile("{\n\t.reg .b64 x, y;\n\tmov.b64 x, {%0, %1};\n\tmov.b64 y, {%2, %2};\n\t"
                 "add.rn.f32x2 x, x, y;\n\tmov.b64 {%0, %1}, x;\n\t}" : "+f"(a0), "+f"(a1) : "f"(b));
}
DEV void op_fmul2(float& a0, float& a1, float b) {
    asm volatile("{\n\t.reg .b64 x, y;\n\tmov.b64 x, {%0, %1};\n\tmov.b64 y, {%2, %2};\n\t"
                 "mul.rn.f32x2 x, x, y;\n\tmov.b64 {%0, %1}, x;\n\t}" : "+f"(a0), "+f"(a1) : "f"(b));
}
DEV void op_max3(float& a, float b, float c) { asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a) : "f"(b), "f"(c)); }
DEV void op_max(float& a, float b) { asm volatile("max.f32 %0, %0, %1;" : "+f"(a) : "f"(b)); }
DEV void op_ex2(float& a) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a)); }
DEV void op_cvt(float& a, float b) {
    uint32_t r;
    asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b));
    a = __uint_as_float(r);
}
DEV void op_imad(float& a, int b) {
    int x = __float_as_int(a);
    asm volatile("mad.lo.s32 %0, %0, 8388608, %1;" : "+r"(x) : "r"(b));
    a = __int_as_float(x);
}
DEV void op_shladd(float& a, int b) {
    int x = __float_as_int(a);
    asm volatile("{\n\t.reg .b32 t;\n\tshl.b32 t, %0, 23;\n\tadd.s32 %0, t, %1;\n\t}" : "+r"(x) : "r"(b));
    a = __int_as_float(x);
}
DEV void op_lop(float& a, int b) {
    int x = __float_as_int(a);
    asm volatile("xor.b32 %0, %0, %1;" : "+r"(x) : "r"(b));
    a = __int_as_float(x);
}
DEV void op_hfma2(float& a, float b) {  // packed bf16 fma on 32-bit registers
    uint32_t x = __float_as_uint(a), y = __float_as_uint(b);
    asm volatile("fma.rn.bf16x2 %0, %0, %1, %1;" : "+r"(x) : "r"(y));
    a = __uint_as_float(x);
}
DEV void op_ex2_h2(float& a) {  // packed half exp2
    uint32_t x = __float_as_uint(a);
    asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(x));
    a = __uint_as_float(x);
}
DEV void op_ex2_f16x2(float& a) {
    uint32_t x = __float_as_uint(a);
    asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(x));
    a = __uint_as_float(x);
}

enum { FFMA, FFMA_IMM, FADD, FFMA2, FADD2, FMUL2, MAX3, MAX2, EX2, CVT, IMAD, SHLADD, LOP, HFMA2, EX2_BF16X2, EX2_F16X2,
       NONE };

template <int OP>
DEV void one(float (&a)[16], int i, float b, float c) {
    const int j = (2 * i) & 15;
    if (OP == FFMA) op_ffma(a[j], b, c);
    if (OP == FFMA_IMM) op_ffma_imm(a[j], c);
    if (OP == FADD) op_fadd(a[j], c);
    if (OP == FFMA2) op_ffma2(a[j], a[j + 1], b, c);
    if (OP == FADD2) op_fadd2(a[j], a[j + 1], c);
    if (OP == FMUL2) op_fmul2(a[j], a[j + 1], b);
    if (OP == MAX3) op_max3(a[j], b, c);
    if (OP == MAX2) op_max(a[j], c);
    if (OP == EX2) op_ex2(a[j]);
    if (OP == CVT) op_cvt(a[j], c);
    if (OP == IMAD) op_imad(a[j], __float_as_int(c));
    if (OP == SHLADD) op_shladd(a[j], __float_as_int(c));
    if (OP == LOP) op_lop(a[j], __float_as_int(c));
    if (OP == HFMA2) op_hfma2(a[j], c);
    if (OP == EX2_BF16X2) op_ex2_h2(a[j]);
    if (OP == EX2_F16X2) op_ex2_f16x2(a[j]);
}

// NA instructions of OPA interleaved with NB of OPB per group, 64 groups per loop iteration
template <int OPA, int NA, int OPB, int NB>
__global__ void __launch_bounds__(512, 1) k(float* out, const float* in, int iters, long long* cyc) {
    float a[16], b2[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        a[i] = in[threadIdx.x % 64 + i];
        b2[i] = in[threadIdx.x % 32 + i + 100];
    }
    const float b = in[1], c = in[2];
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int g = 0; g < 64; ++g) {
#pragma unroll
            for (int x = 0; x < NA; ++x) one<OPA>(a, g * NA + x, b, c);
#pragma unroll
            for (int x = 0; x < NB; ++x) one<OPB>(b2, g * NB + x, b, c);
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    float acc = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc += a[i] + b2[i];
    out[threadIdx.x] = acc;
}

template <int OPA, int NA, int OPB, int NB>
void run(const char* name, float* out, float* in, long long* cyc) {
    printf("%-44s", name);
    for (int wps = 1; wps <= 4; wps *= 2) {
        const int iters = 100;
        k<OPA, NA, OPB, NB><<<1, wps * 128>>>(out, in, iters, cyc);
        cudaDeviceSynchronize();
        k<OPA, NA, OPB, NB><<<1, wps * 128>>>(out, in, iters, cyc);
        cudaDeviceSynchronize();
        long long c;
        cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        // wall cycles per group for each warp, and the scheduler's cost per group (wall / warps per scheduler)
        printf("  w/s %d: wall %6.2f  cost %6.2f |", wps, (double)c / iters / 64, (double)c / iters / 64 / wps);
    }
    printf("\n");
}

int main() {
    float *out, *in;
    long long* cyc;
    cudaMalloc(&out, 4096 * 4);
    cudaMalloc(&in, 4096 * 4);
    cudaMalloc(&cyc, 8);
    float h[4096];
    for (int i = 0; i < 4096; ++i) h[i] = -0.001f * (i % 977);
    h[1] = 0.999f;
    h[2] = -0.0005f;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    printf("group = the listed instructions once; wall = cycles per group seen by a warp; cost = wall / warps per scheduler\n");
    run<FFMA, 1, NONE, 0>("FFMA", out, in, cyc);
    run<FFMA_IMM, 1, NONE, 0>("FFMA imm multiplier", out, in, cyc);
    run<FADD, 1, NONE, 0>("FADD", out, in, cyc);
    run<FFMA2, 1, NONE, 0>("FFMA2", out, in, cyc);
    run<FADD2, 1, NONE, 0>("FADD2", out, in, cyc);
    run<FMUL2, 1, NONE, 0>("FMUL2", out, in, cyc);
    run<MAX3, 1, NONE, 0>("FMNMX3", out, in, cyc);
    run<MAX2, 1, NONE, 0>("FMNMX", out, in, cyc);
    run<EX2, 1, NONE, 0>("MUFU.EX2", out, in, cyc);
    run<CVT, 1, NONE, 0>("F2FP.BF16.PACK_AB", out, in, cyc);
    run<IMAD, 1, NONE, 0>("IMAD (x * 2^23 + y)", out, in, cyc);
    run<SHLADD, 1, NONE, 0>("SHL+IADD (LEA)", out, in, cyc);
    run<LOP, 1, NONE, 0>("LOP3", out, in, cyc);
    run<HFMA2, 1, NONE, 0>("HFMA2.BF16", out, in, cyc);
    run<EX2_BF16X2, 1, NONE, 0>("ex2.bf16x2", out, in, cyc);
    run<EX2_F16X2, 1, NONE, 0>("ex2.f16x2", out, in, cyc);
    run<EX2, 1, FFMA2, 1>("MUFU + FFMA2", out, in, cyc);
    run<EX2, 1, FFMA2, 2>("MUFU + 2 FFMA2", out, in, cyc);
    run<EX2, 1, FFMA2, 4>("MUFU + 4 FFMA2", out, in, cyc);
    run<EX2, 1, FFMA, 4>("MUFU + 4 FFMA", out, in, cyc);
    run<EX2, 1, FFMA, 8>("MUFU + 8 FFMA", out, in, cyc);
    run<EX2, 1, CVT, 1>("MUFU + F2FP", out, in, cyc);
    run<EX2, 1, CVT, 2>("MUFU + 2 F2FP", out, in, cyc);
    run<EX2, 1, MAX3, 2>("MUFU + 2 FMNMX3", out, in, cyc);
    run<EX2, 1, MAX3, 4>("MUFU + 4 FMNMX3", out, in, cyc);
    run<FFMA2, 1, MAX3, 1>("FFMA2 + FMNMX3", out, in, cyc);
    run<FFMA2, 1, CVT, 1>("FFMA2 + F2FP", out, in, cyc);
    run<FFMA2, 1, FFMA, 1>("FFMA2 + FFMA", out, in, cyc);
    run<FFMA2, 1, FADD2, 1>("FFMA2 + FADD2", out, in, cyc);
    run<FFMA, 1, MAX3, 1>("FFMA + FMNMX3", out, in, cyc);
    run<FFMA, 1, CVT, 1>("FFMA + F2FP", out, in, cyc);
    run<FFMA, 1, IMAD, 1>("FFMA + IMAD", out, in, cyc);
    run<CVT, 1, MAX3, 1>("F2FP + FMNMX3", out, in, cyc);
    run<FFMA2, 1, HFMA2, 1>("FFMA2 + HFMA2", out, in, cyc);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
