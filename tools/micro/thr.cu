// Microbenchmarks: per-SM issue throughput of candidate ops for the lattice fast path (one CTA of 32 warps on one SM).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ double rsq(double x){ double r; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); return r; }
__device__ __forceinline__ double rcpa(double x){ double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); return r; }
template<int OP>
__global__ void k(double* out, long long* cyc, int iters, double a, double b, unsigned seed) {
    constexpr int ILP = 8;
    double v[ILP]; unsigned u[ILP]; unsigned long long w[ILP];
    for (int i=0;i<ILP;++i) { v[i] = a + i + threadIdx.x*1e-3; u[i] = seed + i*77u + threadIdx.x; w[i] = u[i]; }
    __syncthreads();
    long long t0 = clock64();
    for (int it=0; it<iters; ++it) {
#pragma unroll
        for (int i=0;i<ILP;++i) {
            if (OP==0) v[i] = __fma_rn(v[i], a, b);
            else if (OP==1) v[i] = __dadd_rn(v[i], b);
            else if (OP==2) { double t; asm volatile("cvt.rn.f64.u32 %0, %1;" : "=d"(t) : "r"(u[i])); u[i] = __double2hiint(t) ^ u[i]; v[i] = t; }
            else if (OP==3) { double t; unsigned short h = (unsigned short)(u[i] >> 16); asm volatile("cvt.rn.f64.u16 %0, %1;" : "=d"(t) : "h"(h)); u[i] = __double2hiint(t) + u[i]; v[i] = t; }
            else if (OP==4) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(u[i]), "r"(seed)); }
            else if (OP==5) v[i] = rsq(v[i]);
            else if (OP==6) v[i] = rcpa(v[i]);
            else if (OP==7) { // DFMA + cvt interleaved: same pipe or not?
                v[i] = __fma_rn(v[i], a, b);
                double t; asm volatile("cvt.rn.f64.u32 %0, %1;" : "=d"(t) : "r"(u[i])); u[i] = __double2hiint(t) ^ u[i];
            }
            else if (OP==8) { // DFMA + MUFU.RSQ64H interleaved
                v[i] = __fma_rn(v[i], a, b);
                double t = rsq(__hiloint2double(u[i] | 0x40000000, 0)); u[i] = __double2hiint(t) ^ u[i];
            }
            else if (OP==9) { // magic-number u2d: LOP + MOV + DADD
                double t = __hiloint2double(0x43300000, (int)(u[i] & 0xffffu)) - 4503599627370496.0; u[i] = __double2loint(t) + u[i] + 1; v[i] = t;
            }
            else if (OP==10) { unsigned m = __match_any_sync(0xffffffffu, u[i] & 7u); u[i] += m; }
            else if (OP==12) { float f = __int_as_float(u[i] | 0x3f800000u); f = __fmaf_rn(f, 1.0000001f, 1e-9f); u[i] = __float_as_int(f) & 0x007fffffu; }   // FFMA chain
            else if (OP==13) { float f; asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(f) : "r"(u[i] & 0xffffu)); u[i] = __float_as_int(f) ^ u[i]; }        // I2FP.F32.U32 (+LOP)
            else if (OP==14) { float f; asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(f) : "f"(__int_as_float(u[i] | 0x40000000u))); u[i] = __float_as_int(f) ^ u[i]; }  // MUFU.RSQ
            else if (OP==15) { float f = __int_as_float(u[i] | 0x40000000u); double t = (double)f; u[i] = __double2hiint(t) ^ u[i]; v[i] = t; }           // F2F.F64.F32
            else if (OP==16) { float f = (float)v[i]; u[i] = __float_as_int(f) ^ u[i]; v[i] = __hiloint2double((int)(u[i] | 0x40000000u) & 0x4fffffff, 0); }   // F2F.F32.F64
            else if (OP==17) { double t = __hiloint2double((int)__byte_perm(u[i], 0x41300000u, 0x7610), 0); v[i] = __fma_rn(t, a, v[i]); u[i] += 3; }     // bias trick + DFMA
            else if (OP==18) { double t; unsigned short h = (unsigned short)(u[i]); asm volatile("cvt.rn.f64.u16 %0, %1;" : "=d"(t) : "h"(h)); v[i] = __fma_rn(t, a, v[i]); u[i] += 3; }  // I2F.F64.U16 + DFMA
            else if (OP==19) { float f = __int_as_float(u[i] | 0x40000000u); float r; asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(f));
                               f = __fmaf_rn(r, f, 1.0f); f = __fmaf_rn(r, f, 1.0f); f = __fmaf_rn(r, f, 1.0f); f = __fmaf_rn(r, f, 1.0f); u[i] = __float_as_int(f) & 0x3fffffffu; }   // MUFU.RSQ + 4 FFMA
            else if (OP==20) { int q; asm volatile("cvt.rni.s32.f32 %0, %1;" : "=r"(q) : "f"(__int_as_float((u[i] & 0x007fffffu) | 0x42000000u))); u[i] += q; }   // F2I
            else if (OP==11) { float f = __uint2float_rn(u[i] & 0xffffu); double t = (double)f; u[i] = __double2hiint(t) ^ u[i]; v[i] = t; }  // I2F.F32 + F2F.F64.F32
        }
    }
    long long t1 = clock64();
    double s=0; for (int i=0;i<ILP;++i) s+=v[i] + u[i] + (double)w[i];
    if (threadIdx.x==0 && blockIdx.x==0) cyc[0] = t1-t0;
    if (s==1.2345) out[0]=s;
}
template<int OP> void run(const char* name, int warps, int ops_per_iter) {
    double* out; long long* cyc; cudaMalloc(&out,8); cudaMalloc(&cyc,8);
    int iters=2048;
    k<OP><<<1,warps*32>>>(out,cyc,iters,1.0000001,1e-9,12345u); cudaDeviceSynchronize();
    k<OP><<<1,warps*32>>>(out,cyc,iters,1.0000001,1e-9,12345u); cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c,cyc,8,cudaMemcpyDeviceToHost);
    double per = (double)c / ((double)iters * 8 * ops_per_iter * warps / 4.0);
    printf("%-28s warps=%2d: %.2f cycles per warp-op per SMSP  (%s)\n", name, warps, per, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}
int main(){
    for (int w : {4, 32}) {
        run<0>("DFMA", w, 1); run<1>("DADD", w, 1); run<2>("cvt.f64.u32 (+LOP)", w, 1); run<3>("cvt.f64.u16 (+shift,add)", w, 1);
        run<4>("IMAD.WIDE.U32", w, 1); run<5>("MUFU.RSQ64H(+mov)", w, 1); run<6>("MUFU.RCP64H(+mov)", w, 1);
        run<7>("DFMA+cvt.f64.u32 pair", w, 1); run<8>("DFMA+RSQ64H pair", w, 1); run<9>("magic u2d (LOP+MOV+DADD)", w, 1);
        run<10>("MATCH.ANY", w, 1); run<11>("I2F.F32+F2F.F64.F32", w, 1);
        run<12>("FFMA (+2 LOP)", w, 1); run<13>("I2FP.F32.U32 (+2 LOP)", w, 1); run<14>("MUFU.RSQ f32 (+2 LOP)", w, 1);
        run<15>("F2F.F64.F32 (+LOP)", w, 1); run<16>("F2F.F32.F64 (+LOPs)", w, 1); run<17>("PRMT bias + DFMA", w, 1);
        run<18>("I2F.F64.U16 + DFMA", w, 1); run<19>("MUFU.RSQ + 4 FFMA", w, 1); run<20>("F2I.S32.F32 (+LOPs)", w, 1);
    }
    return 0;
}
