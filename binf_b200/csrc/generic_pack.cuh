// binfb_f2: two chains side by side in one register pair.  A user-defined forward model (generic_kernel.cuh)
// written over `float` is compiled a second time with `float` standing for this type, so that its
// arithmetic is issued as packed FP32 (FFMA2 / FADD2 / FMUL2: one instruction for the two chains of a
// lane).  Every operation acts on the two components independently and rounds like its scalar counterpart.
// Part of the NVRTC source bundle (build.py), not compiled by nvcc.
struct binfb_f2 {
    float2 v;
    __device__ __forceinline__ binfb_f2() {}
    __device__ __forceinline__ binfb_f2(float a) : v(make_float2(a, a)) {}
    __device__ __forceinline__ binfb_f2(double a) : v(make_float2((float)a, (float)a)) {}
    __device__ __forceinline__ binfb_f2(int a) : v(make_float2((float)a, (float)a)) {}
    __device__ __forceinline__ binfb_f2(float a, float b) : v(make_float2(a, b)) {}
    __device__ __forceinline__ explicit binfb_f2(float2 a) : v(a) {}
};
__device__ __forceinline__ binfb_f2 operator+(binfb_f2 a, binfb_f2 b) { return binfb_f2(__fadd2_rn(a.v, b.v)); }
__device__ __forceinline__ binfb_f2 operator-(binfb_f2 a, binfb_f2 b) {
    return binfb_f2(__fadd2_rn(a.v, make_float2(-b.v.x, -b.v.y)));
}
__device__ __forceinline__ binfb_f2 operator*(binfb_f2 a, binfb_f2 b) { return binfb_f2(__fmul2_rn(a.v, b.v)); }
__device__ __forceinline__ binfb_f2 operator/(binfb_f2 a, binfb_f2 b) { return binfb_f2(a.v.x / b.v.x, a.v.y / b.v.y); }
__device__ __forceinline__ binfb_f2 operator-(binfb_f2 a) { return binfb_f2(-a.v.x, -a.v.y); }
__device__ __forceinline__ binfb_f2 operator+(binfb_f2 a) { return a; }
__device__ __forceinline__ binfb_f2 &operator+=(binfb_f2 &a, binfb_f2 b) { return a = a + b; }
__device__ __forceinline__ binfb_f2 &operator-=(binfb_f2 &a, binfb_f2 b) { return a = a - b; }
__device__ __forceinline__ binfb_f2 &operator*=(binfb_f2 &a, binfb_f2 b) { return a = a * b; }
__device__ __forceinline__ binfb_f2 &operator/=(binfb_f2 &a, binfb_f2 b) { return a = a / b; }
__device__ __forceinline__ binfb_f2 fmaf(binfb_f2 a, binfb_f2 b, binfb_f2 c) { return binfb_f2(__ffma2_rn(a.v, b.v, c.v)); }
__device__ __forceinline__ binfb_f2 __fmaf_rn(binfb_f2 a, binfb_f2 b, binfb_f2 c) { return binfb_f2(__ffma2_rn(a.v, b.v, c.v)); }
#define BINFB_F2_UNARY(fn) \
    __device__ __forceinline__ binfb_f2 fn(binfb_f2 a) { return binfb_f2(fn(a.v.x), fn(a.v.y)); }
#define BINFB_F2_BINARY(fn) \
    __device__ __forceinline__ binfb_f2 fn(binfb_f2 a, binfb_f2 b) { return binfb_f2(fn(a.v.x, b.v.x), fn(a.v.y, b.v.y)); }
BINFB_F2_UNARY(sqrtf) BINFB_F2_UNARY(rsqrtf) BINFB_F2_UNARY(expf) BINFB_F2_UNARY(exp2f) BINFB_F2_UNARY(expm1f)
BINFB_F2_UNARY(logf) BINFB_F2_UNARY(log2f) BINFB_F2_UNARY(log1pf) BINFB_F2_UNARY(sinf) BINFB_F2_UNARY(cosf)
BINFB_F2_UNARY(tanf) BINFB_F2_UNARY(tanhf) BINFB_F2_UNARY(sinhf) BINFB_F2_UNARY(coshf) BINFB_F2_UNARY(atanf)
BINFB_F2_UNARY(fabsf) BINFB_F2_UNARY(erff) BINFB_F2_UNARY(erfcf) BINFB_F2_UNARY(__expf) BINFB_F2_UNARY(__logf)
BINFB_F2_UNARY(__sinf) BINFB_F2_UNARY(__cosf) BINFB_F2_UNARY(__frcp_rn) BINFB_F2_UNARY(__frsqrt_rn)
BINFB_F2_BINARY(fminf) BINFB_F2_BINARY(fmaxf) BINFB_F2_BINARY(powf) BINFB_F2_BINARY(atan2f) BINFB_F2_BINARY(__fdividef)
BINFB_F2_BINARY(copysignf)
#undef BINFB_F2_UNARY
#undef BINFB_F2_BINARY
