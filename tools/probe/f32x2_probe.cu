// Probe (no GPU needed): can the issue-bound SOR kernels use Blackwell's packed fp32 instructions (PTX add/mul/fma
// .f32x2 -> SASS FADD2/FMUL2/FFMA2: two IEEE fp32 operations per issue slot; (du,dv) are natural pairs)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -cubin -o /tmp/f2.cubin tools/probe/f32x2_probe.cu
//   cuobjdump -sass /tmp/f2.cubin | grep -E "FMUL2|FADD2|FFMA2"
// Finding (CUDA 12.9.86): ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 although both carry an explicit
// rounding modifier and -fmad=false is given to nvcc and to ptxas (the scalar forms are never contracted then):
//   FMUL2 R12, R12.F32x2.HI_LO, UR7.F32 ; FFMA2 R12, R4.F32x2.HI_LO, UR6.F32, R12.F32x2.HI_LO
// A fused product is rounded once, the reference's SSE build rounds twice: not usable for the bit-exact path as is.
// The all-FMA spelling -- product = fma.rn.f32x2(a, b, -0), sum = fma.rn.f32x2(p, 1, q), each exactly one rounding -- is
// canonicalised back to mul/add and contracted the same way.
// (The broadcast operand form `UR7.F32` shows that a scalar times a pair needs no packing instruction.)
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pk(float a, float b){unsigned long long r; asm("mov.b64 %0, {%1,%2};":"=l"(r):"f"(a),"f"(b)); return r;}
__device__ __forceinline__ void upk(unsigned long long v, float&a, float&b){asm("mov.b64 {%0,%1}, %2;":"=f"(a),"=f"(b):"l"(v));}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b){unsigned long long r; asm("mul.rn.f32x2 %0, %1, %2;":"=l"(r):"l"(a),"l"(b)); return r;}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b){unsigned long long r; asm("add.rn.f32x2 %0, %1, %2;":"=l"(r):"l"(a),"l"(b)); return r;}
__global__ void k(const float4* in, float4* out, float hh, float vt){
  float4 a=in[threadIdx.x], b=in[threadIdx.x+32];
  unsigned long long x=pk(a.x,a.y), y=pk(b.x,b.y), h2=pk(hh,hh), v2=pk(vt,vt);
  unsigned long long t=add2(mul2(h2,x), mul2(v2,y));
  unsigned long long u=add2(mul2(h2,pk(a.z,a.w)), mul2(v2,pk(b.z,b.w)));
  float4 o; upk(t,o.x,o.y); upk(u,o.z,o.w);
  out[threadIdx.x]=o;
}
