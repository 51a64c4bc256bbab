// Micro-benchmark: per-SMSP issue throughput of the ALU ops the traversal kernel is built from (sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_alu ubench_alu.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

#define ITER 256
#define CHAINS 8

template<int OP>
__global__ void __launch_bounds__(1024) k(uint32_t *out, uint32_t seed, long long *cycles) {
    uint32_t r[CHAINS];
    float f[CHAINS];
    unsigned long long p[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) r[i] = seed * (i + 3) + threadIdx.x, f[i] = (float)(seed + i) * 1.0001f, p[i] = ((unsigned long long)r[i] << 32) | (seed + i);
    float fa = __uint_as_float(0x3f7ffff0u + seed), fb = (float)seed * 1e-9f;
    unsigned long long pa = ((unsigned long long)__float_as_uint(fa) << 32) | __float_as_uint(fa);
    unsigned long long pb = ((unsigned long long)__float_as_uint(fb) << 32) | __float_as_uint(fb);
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int i = 0; i < CHAINS; i++) {
                if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fa), "f"(fb));
                if (OP == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));
                if (OP == 2) { asm volatile("cvt.rn.f32.u8 %0, %1;" : "=f"(f[i]) : "r"(r[i])); r[i] = __float_as_uint(f[i]); }
                if (OP == 3) asm volatile("prmt.b32 %0, %0, %1, 0x7144;" : "+r"(r[i]) : "r"(seed));
                if (OP == 4) asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fa));
                if (OP == 5) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fa), "f"(fb));
                if (OP == 6) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(seed), "r"(r[(i + 1) % CHAINS]));
                if (OP == 7) asm volatile("{.reg .pred q; setp.le.f32 q, %0, %1; selp.f32 %0, %2, %0, q;}" : "+f"(f[i]) : "f"(fa), "f"(fb));
                if (OP == 8) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(seed), "r"(seed));
                if (OP == 9) asm volatile("shl.b32 %0, %0, %1;" : "+r"(r[i]) : "r"(seed & 1));
                if (OP == 10) asm volatile("add.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fb));
                if (OP == 11) asm volatile("{.reg .pred q; setp.le.f32 q, %1, %2; @q or.b32 %0, %0, %3;}" : "+r"(r[i]) : "f"(f[i]), "f"(fa), "r"(seed));
                if (OP == 12) asm volatile("add.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(seed));
                if (OP == 13) asm volatile("{.reg .f64 d; mov.b64 d, %0; fma.rn.f64 d, d, d, d; mov.b64 %0, d;}" : "+l"(p[i]));
                if (OP == 14) asm volatile("bfe.u32 %0, %0, 8, 8;" : "+r"(r[i]));
                if (OP == 15) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
                if (OP == 16) { asm volatile("{.reg .b16 lo, hi; mov.b32 {lo, hi}, %1; cvt.f32.f16 %0, hi;}" : "=f"(f[i]) : "r"(r[i])); r[i] = __float_as_uint(f[i]); }
                if (OP == 17) { asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(f[i]) : "r"(r[i])); r[i] = __float_as_uint(f[i]); }
                if (OP == 18) { asm volatile("{.reg .b16 lo, hi; mov.b32 {lo, hi}, %1; cvt.rn.f32.u16 %0, lo;}" : "=f"(f[i]) : "r"(r[i])); r[i] = __float_as_uint(f[i]); }
                if (OP == 19) { asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fa)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(seed), "r"(r[(i + 1) % CHAINS])); }
                if (OP == 20) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fa), "f"(fb)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(seed), "r"(r[(i + 1) % CHAINS])); }
                if (OP == 21) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fa), "f"(fb)); asm volatile("max.f32 %0, %0, %1;" : "+f"(f[(i + 4) % CHAINS]) : "f"(fa)); }
                if (OP == 22) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fa), "f"(fb)); asm volatile("prmt.b32 %0, %0, %1, 0x7144;" : "+r"(r[i]) : "r"(seed)); }
                if (OP == 23) { asm volatile("add.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(seed)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[(i+4)%CHAINS]) : "r"(seed), "r"(r[(i + 1) % CHAINS])); }
                if (OP == 24) asm volatile("{.reg .pred q; setp.le.f32 q, %1, %2; @q add.u32 %0, %0, 16;}" : "+r"(r[i]) : "f"(f[i]), "f"(fa));
                if (OP == 25) { asm volatile("{.reg .b16 lo, hi; mov.b32 {lo, hi}, %1; cvt.f32.f16 %0, hi;}" : "=f"(f[i]) : "r"(r[i])); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fa), "f"(fb)); r[i] = __float_as_uint(f[i]); }
                if (OP == 26) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb)); asm volatile("prmt.b32 %0, %0, %1, 0x7144;" : "+r"(r[i]) : "r"(seed)); asm volatile("prmt.b32 %0, %0, %1, 0x7144;" : "+r"(r[(i+4)%CHAINS]) : "r"(seed));}
                if (OP == 27) asm volatile("{.reg .pred q; setp.le.f32 q, %0, %1; selp.f32 %0, %2, %0, q;}" : "+f"(f[i]) : "f"(fa), "f"(fb));
                if (OP == 28) asm volatile("{.reg .b32 t; bfind.u32 t, %0; add.u32 %0, %0, t;}" : "+r"(r[i]));
                if (OP == 29) asm volatile("{.reg .b32 t; popc.b32 t, %0; add.u32 %0, %0, t;}" : "+r"(r[i]));
                if (OP == 31) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(seed), "r"(seed));
                if (OP == 32) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fa), "f"(fb)); asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(seed), "r"(seed)); }
                if (OP == 33) { asm volatile("prmt.b32 %0, %0, %1, 0x7144;" : "+r"(r[i]) : "r"(seed)); asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(r[(i+4)%CHAINS]) : "r"(seed), "r"(seed)); }
                if (OP == 34) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb)); asm volatile("prmt.b32 %0, %0, %1, 0x7144;" : "+r"(r[i]) : "r"(seed)); asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(r[(i+4)%CHAINS]) : "r"(seed), "r"(seed));}
                if (OP == 35) { asm volatile("prmt.b32 %0, %0, %1, 0x7144;" : "+r"(r[i]) : "r"(seed)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[(i+4)%CHAINS]) : "r"(seed), "r"(seed)); }
                if (OP == 36) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fa), "f"(fb)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(seed), "r"(seed)); }
                if (OP == 37) { asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fa)); asm volatile("prmt.b32 %0, %0, %1, 0x7144;" : "+r"(r[i]) : "r"(seed)); }
                if (OP == 38) { asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fa), "f"(fb)); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[(i+4)%CHAINS]) : "f"(fa), "f"(fb)); }
                if (OP == 39) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb)); asm volatile("prmt.b32 %0, %0, %1, 0x7144;" : "+r"(r[i]) : "r"(seed)); }
                if (OP == 40) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb)); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fa), "f"(fb)); }
                if (OP == 30) asm volatile("{.reg .pred q; setp.le.f32 q, %0, %1; @q mov.f32 %0, %2;}" : "+f"(f[i]) : "f"(fa), "f"(fb));
            }
        }
    }
    long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) acc ^= r[i] ^ __float_as_uint(f[i]) ^ (uint32_t)p[i] ^ (uint32_t)(p[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template<int OP>
void run(const char *name, int threads) {
    uint32_t *out;
    long long *cyc;
    int blocks = 148;
    cudaMalloc(&out, blocks * 1024 * 4);
    cudaMalloc(&cyc, blocks * 8);
    k<OP><<<blocks, threads>>>(out, 1, cyc);
    k<OP><<<blocks, threads>>>(out, 1, cyc);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < blocks; i++) avg += h[i];
    avg /= blocks;
    double warp_instr_per_smsp = (double)ITER * 4 * CHAINS * (threads / 32) / 4.0;
    printf("%-28s threads/SM=%4d  cycles=%9.0f  cycles per warp-instr per SMSP = %.3f\n", name, threads, avg, avg / warp_instr_per_smsp);
    cudaFree(out), cudaFree(cyc);
}

int main() {
    for (int threads : {1024}) {
        run<0>("FFMA", threads);
        run<1>("FFMA2 (f32x2)", threads);
        run<2>("I2F.U8 (cvt.rn.f32.u8)", threads);
        run<3>("PRMT", threads);
        run<4>("FMNMX", threads);
        run<5>("FMNMX3", threads);
        run<6>("LOP3", threads);
        run<7>("FSETP+FSEL", threads);
        run<8>("IMAD", threads);
        run<9>("SHL", threads);
        run<10>("FADD", threads);
        run<11>("FSETP+@p LOP", threads);
        run<12>("IADD", threads);
        run<13>("DFMA", threads);
        run<14>("BFE", threads);
        run<15>("MUFU.RCP", threads);
        run<16>("cvt.f32.f16 (HADD2.F32?)", threads);
        run<17>("cvt.rn.f32.u32 (I2FP)", threads);
        run<18>("cvt.rn.f32.u16", threads);
        run<19>("FMNMX + LOP3 pair", threads);
        run<20>("FFMA + LOP3 pair", threads);
        run<21>("FFMA + FMNMX pair", threads);
        run<22>("FFMA + PRMT pair", threads);
        run<23>("IADD + LOP3 pair", threads);
        run<24>("FSETP + @p IADD", threads);
        run<25>("cvt.f32.f16 + FFMA pair", threads);
        run<26>("FFMA2 + 2 PRMT triple", threads);
        run<28>("BFIND+IADD", threads);
        run<29>("POPC+IADD", threads);
        run<30>("FSETP+@p MOV", threads);
        run<31>("IDP4A", threads);
        run<32>("FFMA + IDP4A pair", threads);
        run<33>("PRMT + IDP4A pair", threads);
        run<34>("FFMA2 + PRMT + IDP4A triple", threads);
        run<35>("PRMT + IMAD pair", threads);
        run<36>("FFMA + IMAD pair", threads);
        run<37>("FMNMX + PRMT pair", threads);
        run<38>("FMNMX3 + FFMA pair", threads);
        run<39>("FFMA2 + PRMT pair", threads);
        run<40>("FFMA2 + FFMA pair", threads);
    }
    return 0;
}
