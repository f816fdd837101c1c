// wgb_prelude.cuh -- device-side WGSL value types and operators used by emitted shader code.
//
// Arithmetic contract: wgpu-cpu's shader JIT emits one IEEE-754 binary32
// operation per WGSL operator and never fuses a multiply with an add
// (naga-cranelift/src/expression/binary.rs:420-449).  Every float operator below is
// therefore spelled with the round-to-nearest intrinsics (__fmul_rn, __fadd_rn, ...),
// which the compiler may not contract into FMAs, and the pipeline TU is additionally
// compiled with --fmad=false, default --prec-div=true / --ftz=false.
#pragma once
#include "wgb_shared.h"

typedef unsigned int u32;
typedef int i32;
typedef float f32;

#define WGB_DEV __device__ __forceinline__

WGB_DEV f32 wgb_add(f32 a, f32 b) { return __fadd_rn(a, b); }
WGB_DEV f32 wgb_sub(f32 a, f32 b) { return __fsub_rn(a, b); }
WGB_DEV f32 wgb_mul(f32 a, f32 b) { return __fmul_rn(a, b); }
WGB_DEV f32 wgb_div(f32 a, f32 b) { return __fdiv_rn(a, b); }
// float % : a - b*trunc(a/b) as four separate ops (binary.rs:435-449)
WGB_DEV f32 wgb_rem(f32 a, f32 b) { return __fsub_rn(a, __fmul_rn(b, truncf(__fdiv_rn(a, b)))); }

// ---------------------------------------------------------------------------------------
// f16 (`enable f16;` -- the scalar exists in the reference's type table, naga-cranelift/src/types.rs:103-136).
// A value is its IEEE binary16 bits; every operator converts to binary32 (exact), applies ONE binary32 operation and
// rounds back to binary16.  binary32 carries 24 >= 2 * 11 + 2 significand bits, so for + - * / the second rounding
// cannot disagree with rounding the exact result once: these are the correctly rounded binary16 operations.
// ---------------------------------------------------------------------------------------
#ifndef WGB_F16_CONVERSIONS_PROVIDED     // (tests/cusim converts through the host compiler's _Float16)
WGB_DEV unsigned short wgb_f32_to_f16_bits(f32 v) { unsigned short h; asm("cvt.rn.f16.f32 %0, %1;" : "=h"(h) : "f"(v)); return h; }
WGB_DEV f32 wgb_f16_bits_to_f32(unsigned short h) { f32 v; asm("cvt.f32.f16 %0, %1;" : "=f"(v) : "h"(h)); return v; }
#endif
struct f16 {
    unsigned short bits;
    WGB_DEV f16() : bits(0) {}
    WGB_DEV explicit f16(int) : bits(0) {}            // f16(0): the zero value the emitter writes for an uninitialised var
};
WGB_DEV f16 wgb_f16(f32 v) { f16 h; h.bits = wgb_f32_to_f16_bits(v); return h; }
WGB_DEV f32 wgb_f32(f16 h) { return wgb_f16_bits_to_f32(h.bits); }
WGB_DEV f16 wgb_add(f16 a, f16 b) { return wgb_f16(__fadd_rn(wgb_f32(a), wgb_f32(b))); }
WGB_DEV f16 wgb_sub(f16 a, f16 b) { return wgb_f16(__fsub_rn(wgb_f32(a), wgb_f32(b))); }
WGB_DEV f16 wgb_mul(f16 a, f16 b) { return wgb_f16(__fmul_rn(wgb_f32(a), wgb_f32(b))); }
WGB_DEV f16 wgb_div(f16 a, f16 b) { return wgb_f16(__fdiv_rn(wgb_f32(a), wgb_f32(b))); }
// a - b*trunc(a/b), every step a binary16 operation (as wgb_rem for f32)
WGB_DEV f16 wgb_rem(f16 a, f16 b) { return wgb_sub(a, wgb_mul(b, wgb_f16(truncf(wgb_f32(wgb_div(a, b)))))); }
WGB_DEV f16 operator-(f16 a) { f16 r; r.bits = (unsigned short)(a.bits ^ 0x8000u); return r; }
WGB_DEV f16 operator+(f16 a, f16 b) { return wgb_add(a, b); }
WGB_DEV f16 operator-(f16 a, f16 b) { return wgb_sub(a, b); }
WGB_DEV f16 operator*(f16 a, f16 b) { return wgb_mul(a, b); }
WGB_DEV f16 operator/(f16 a, f16 b) { return wgb_div(a, b); }
WGB_DEV bool operator==(f16 a, f16 b) { return wgb_f32(a) == wgb_f32(b); }
WGB_DEV bool operator!=(f16 a, f16 b) { return wgb_f32(a) != wgb_f32(b); }
WGB_DEV bool operator<(f16 a, f16 b) { return wgb_f32(a) < wgb_f32(b); }
WGB_DEV bool operator>(f16 a, f16 b) { return wgb_f32(a) > wgb_f32(b); }
WGB_DEV bool operator<=(f16 a, f16 b) { return wgb_f32(a) <= wgb_f32(b); }
WGB_DEV bool operator>=(f16 a, f16 b) { return wgb_f32(a) >= wgb_f32(b); }

// ---------------------------------------------------------------------------------------
// vectors
// ---------------------------------------------------------------------------------------
#define WGB_VEC_TYPES(T, S)                                                                         \
    struct vec2##S {                                                                                \
        T x, y;                                                                                     \
        WGB_DEV vec2##S() : x(0), y(0) {}                                                           \
        WGB_DEV vec2##S(T x_, T y_) : x(x_), y(y_) {}                                               \
        WGB_DEV explicit vec2##S(T s) : x(s), y(s) {}                                               \
        WGB_DEV T& operator[](int i) { return (&x)[i]; }                                            \
        WGB_DEV const T& operator[](int i) const { return (&x)[i]; }                                \
    };                                                                                              \
    struct vec3##S {                                                                                \
        T x, y, z;                                                                                  \
        WGB_DEV vec3##S() : x(0), y(0), z(0) {}                                                     \
        WGB_DEV vec3##S(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}                                  \
        WGB_DEV explicit vec3##S(T s) : x(s), y(s), z(s) {}                                         \
        WGB_DEV vec3##S(vec2##S a, T z_) : x(a.x), y(a.y), z(z_) {}                                 \
        WGB_DEV vec3##S(T x_, vec2##S a) : x(x_), y(a.x), z(a.y) {}                                 \
        WGB_DEV T& operator[](int i) { return (&x)[i]; }                                            \
        WGB_DEV const T& operator[](int i) const { return (&x)[i]; }                                \
    };                                                                                              \
    struct vec4##S {                                                                                \
        T x, y, z, w;                                                                               \
        WGB_DEV vec4##S() : x(0), y(0), z(0), w(0) {}                                               \
        WGB_DEV vec4##S(T x_, T y_, T z_, T w_) : x(x_), y(y_), z(z_), w(w_) {}                     \
        WGB_DEV explicit vec4##S(T s) : x(s), y(s), z(s), w(s) {}                                   \
        WGB_DEV vec4##S(vec2##S a, T z_, T w_) : x(a.x), y(a.y), z(z_), w(w_) {}                    \
        WGB_DEV vec4##S(T x_, vec2##S a, T w_) : x(x_), y(a.x), z(a.y), w(w_) {}                    \
        WGB_DEV vec4##S(T x_, T y_, vec2##S a) : x(x_), y(y_), z(a.x), w(a.y) {}                    \
        WGB_DEV vec4##S(vec2##S a, vec2##S b) : x(a.x), y(a.y), z(b.x), w(b.y) {}                   \
        WGB_DEV vec4##S(vec3##S a, T w_) : x(a.x), y(a.y), z(a.z), w(w_) {}                         \
        WGB_DEV vec4##S(T x_, vec3##S a) : x(x_), y(a.x), z(a.y), w(a.z) {}                         \
        WGB_DEV T& operator[](int i) { return (&x)[i]; }                                            \
        WGB_DEV const T& operator[](int i) const { return (&x)[i]; }                                \
    };

WGB_VEC_TYPES(f32, f)
WGB_VEC_TYPES(i32, i)
WGB_VEC_TYPES(u32, u)
WGB_VEC_TYPES(bool, b)
WGB_VEC_TYPES(f16, h)

// element-wise float operators; scalar (x) vector in both orders (binary.rs:210-268, 343-418)
#define WGB_VEC_FLOAT_OP(S, T, OP, FN)                                                                            \
    WGB_DEV vec2##S operator OP(vec2##S a, vec2##S b) { return vec2##S(FN(a.x, b.x), FN(a.y, b.y)); }              \
    WGB_DEV vec3##S operator OP(vec3##S a, vec3##S b) { return vec3##S(FN(a.x, b.x), FN(a.y, b.y), FN(a.z, b.z)); } \
    WGB_DEV vec4##S operator OP(vec4##S a, vec4##S b) { return vec4##S(FN(a.x, b.x), FN(a.y, b.y), FN(a.z, b.z), FN(a.w, b.w)); } \
    WGB_DEV vec2##S operator OP(vec2##S a, T s) { return vec2##S(FN(a.x, s), FN(a.y, s)); }                        \
    WGB_DEV vec3##S operator OP(vec3##S a, T s) { return vec3##S(FN(a.x, s), FN(a.y, s), FN(a.z, s)); }            \
    WGB_DEV vec4##S operator OP(vec4##S a, T s) { return vec4##S(FN(a.x, s), FN(a.y, s), FN(a.z, s), FN(a.w, s)); } \
    WGB_DEV vec2##S operator OP(T s, vec2##S a) { return vec2##S(FN(s, a.x), FN(s, a.y)); }                        \
    WGB_DEV vec3##S operator OP(T s, vec3##S a) { return vec3##S(FN(s, a.x), FN(s, a.y), FN(s, a.z)); }            \
    WGB_DEV vec4##S operator OP(T s, vec4##S a) { return vec4##S(FN(s, a.x), FN(s, a.y), FN(s, a.z), FN(s, a.w)); }
WGB_VEC_FLOAT_OP(f, f32, +, wgb_add)
WGB_VEC_FLOAT_OP(f, f32, -, wgb_sub)
WGB_VEC_FLOAT_OP(f, f32, *, wgb_mul)
WGB_VEC_FLOAT_OP(f, f32, /, wgb_div)
WGB_VEC_FLOAT_OP(h, f16, +, wgb_add)
WGB_VEC_FLOAT_OP(h, f16, -, wgb_sub)
WGB_VEC_FLOAT_OP(h, f16, *, wgb_mul)
WGB_VEC_FLOAT_OP(h, f16, /, wgb_div)
WGB_DEV vec2h operator-(vec2h a) { return vec2h(-a.x, -a.y); }
WGB_DEV vec3h operator-(vec3h a) { return vec3h(-a.x, -a.y, -a.z); }
WGB_DEV vec4h operator-(vec4h a) { return vec4h(-a.x, -a.y, -a.z, -a.w); }
WGB_DEV vec2f operator-(vec2f a) { return vec2f(-a.x, -a.y); }
WGB_DEV vec3f operator-(vec3f a) { return vec3f(-a.x, -a.y, -a.z); }
WGB_DEV vec4f operator-(vec4f a) { return vec4f(-a.x, -a.y, -a.z, -a.w); }

#define WGB_VEC_INT_OP(S, T, OP)                                                                                  \
    WGB_DEV vec2##S operator OP(vec2##S a, vec2##S b) { return vec2##S(a.x OP b.x, a.y OP b.y); }                  \
    WGB_DEV vec3##S operator OP(vec3##S a, vec3##S b) { return vec3##S(a.x OP b.x, a.y OP b.y, a.z OP b.z); }      \
    WGB_DEV vec4##S operator OP(vec4##S a, vec4##S b) { return vec4##S(a.x OP b.x, a.y OP b.y, a.z OP b.z, a.w OP b.w); } \
    WGB_DEV vec2##S operator OP(vec2##S a, T s) { return vec2##S(a.x OP s, a.y OP s); }                            \
    WGB_DEV vec3##S operator OP(vec3##S a, T s) { return vec3##S(a.x OP s, a.y OP s, a.z OP s); }                  \
    WGB_DEV vec4##S operator OP(vec4##S a, T s) { return vec4##S(a.x OP s, a.y OP s, a.z OP s, a.w OP s); }
WGB_VEC_INT_OP(i, i32, +)
WGB_VEC_INT_OP(i, i32, -)
WGB_VEC_INT_OP(i, i32, *)
WGB_VEC_INT_OP(u, u32, +)
WGB_VEC_INT_OP(u, u32, -)
WGB_VEC_INT_OP(u, u32, *)

// ---------------------------------------------------------------------------------------
// matrices (column-major, columns are vectors; types.rs:442-444)
// ---------------------------------------------------------------------------------------
struct mat2x2f {
    vec2f c[2];
    WGB_DEV mat2x2f() {}
    WGB_DEV mat2x2f(vec2f c0, vec2f c1) { c[0] = c0; c[1] = c1; }
    WGB_DEV vec2f& operator[](int i) { return c[i]; }
    WGB_DEV const vec2f& operator[](int i) const { return c[i]; }
};
struct mat3x3f {
    vec3f c[3];
    WGB_DEV mat3x3f() {}
    WGB_DEV mat3x3f(vec3f c0, vec3f c1, vec3f c2) { c[0] = c0; c[1] = c1; c[2] = c2; }
    WGB_DEV vec3f& operator[](int i) { return c[i]; }
    WGB_DEV const vec3f& operator[](int i) const { return c[i]; }
};
struct mat4x4f {
    vec4f c[4];
    WGB_DEV mat4x4f() {}
    WGB_DEV mat4x4f(vec4f c0, vec4f c1, vec4f c2, vec4f c3) { c[0] = c0; c[1] = c1; c[2] = c2; c[3] = c3; }
    WGB_DEV vec4f& operator[](int i) { return c[i]; }
    WGB_DEV const vec4f& operator[](int i) const { return c[i]; }
};
// matrix * vector: x_i = splat(v[i]) * M.col[i], accumulated left to right with separate
// adds: ((v0*c0 + v1*c1) + v2*c2) + v3*c3   (binary.rs:297-323)
WGB_DEV vec2f operator*(const mat2x2f& m, vec2f v) { return v.x * m.c[0] + v.y * m.c[1]; }
WGB_DEV vec3f operator*(const mat3x3f& m, vec3f v) { return (v.x * m.c[0] + v.y * m.c[1]) + v.z * m.c[2]; }
WGB_DEV vec4f operator*(const mat4x4f& m, vec4f v) { return ((v.x * m.c[0] + v.y * m.c[1]) + v.z * m.c[2]) + v.w * m.c[3]; }
WGB_DEV mat4x4f operator*(const mat4x4f& m, f32 s) { return mat4x4f(m.c[0] * s, m.c[1] * s, m.c[2] * s, m.c[3] * s); }
WGB_DEV mat4x4f operator*(f32 s, const mat4x4f& m) { return mat4x4f(s * m.c[0], s * m.c[1], s * m.c[2], s * m.c[3]); }
// matrix * matrix is a todo!() in the reference (binary.rs:254); defined column by column with the same rule
WGB_DEV mat4x4f operator*(const mat4x4f& a, const mat4x4f& b) { return mat4x4f(a * b.c[0], a * b.c[1], a * b.c[2], a * b.c[3]); }
WGB_DEV mat3x3f operator*(const mat3x3f& a, const mat3x3f& b) { return mat3x3f(a * b.c[0], a * b.c[1], a * b.c[2]); }
WGB_DEV mat2x2f operator*(const mat2x2f& a, const mat2x2f& b) { return mat2x2f(a * b.c[0], a * b.c[1]); }
WGB_DEV mat3x3f operator*(const mat3x3f& m, f32 s) { return mat3x3f(m.c[0] * s, m.c[1] * s, m.c[2] * s); }
WGB_DEV mat3x3f operator*(f32 s, const mat3x3f& m) { return mat3x3f(s * m.c[0], s * m.c[1], s * m.c[2]); }
WGB_DEV mat2x2f operator*(const mat2x2f& m, f32 s) { return mat2x2f(m.c[0] * s, m.c[1] * s); }
WGB_DEV mat2x2f operator*(f32 s, const mat2x2f& m) { return mat2x2f(s * m.c[0], s * m.c[1]); }
WGB_DEV mat4x4f operator+(const mat4x4f& a, const mat4x4f& b) { return mat4x4f(a.c[0] + b.c[0], a.c[1] + b.c[1], a.c[2] + b.c[2], a.c[3] + b.c[3]); }
WGB_DEV mat4x4f operator-(const mat4x4f& a, const mat4x4f& b) { return mat4x4f(a.c[0] - b.c[0], a.c[1] - b.c[1], a.c[2] - b.c[2], a.c[3] - b.c[3]); }
WGB_DEV mat3x3f operator+(const mat3x3f& a, const mat3x3f& b) { return mat3x3f(a.c[0] + b.c[0], a.c[1] + b.c[1], a.c[2] + b.c[2]); }
WGB_DEV mat3x3f operator-(const mat3x3f& a, const mat3x3f& b) { return mat3x3f(a.c[0] - b.c[0], a.c[1] - b.c[1], a.c[2] - b.c[2]); }
WGB_DEV mat2x2f operator+(const mat2x2f& a, const mat2x2f& b) { return mat2x2f(a.c[0] + b.c[0], a.c[1] + b.c[1]); }
WGB_DEV mat2x2f operator-(const mat2x2f& a, const mat2x2f& b) { return mat2x2f(a.c[0] - b.c[0], a.c[1] - b.c[1]); }
WGB_DEV mat4x4f wgb_transpose(const mat4x4f& m) {
    return mat4x4f(vec4f(m.c[0].x, m.c[1].x, m.c[2].x, m.c[3].x), vec4f(m.c[0].y, m.c[1].y, m.c[2].y, m.c[3].y),
                   vec4f(m.c[0].z, m.c[1].z, m.c[2].z, m.c[3].z), vec4f(m.c[0].w, m.c[1].w, m.c[2].w, m.c[3].w));
}
WGB_DEV mat3x3f wgb_transpose(const mat3x3f& m) {
    return mat3x3f(vec3f(m.c[0].x, m.c[1].x, m.c[2].x), vec3f(m.c[0].y, m.c[1].y, m.c[2].y), vec3f(m.c[0].z, m.c[1].z, m.c[2].z));
}
WGB_DEV mat2x2f wgb_transpose(const mat2x2f& m) { return mat2x2f(vec2f(m.c[0].x, m.c[1].x), vec2f(m.c[0].y, m.c[1].y)); }

// ---------------------------------------------------------------------------------------
// casts (expression/as.rs:60-184).  float -> int traps in the reference when out of
// range; here it saturates (CUDA's cvt.rzi), which agrees wherever the reference runs.
// ---------------------------------------------------------------------------------------
WGB_DEV f32 wgb_to_f32(f32 v) { return v; }
WGB_DEV f32 wgb_to_f32(i32 v) { return __int2float_rn(v); }
WGB_DEV f32 wgb_to_f32(u32 v) { return __uint2float_rn(v); }
WGB_DEV f32 wgb_to_f32(bool v) { return v ? 1.0f : 0.0f; }
WGB_DEV i32 wgb_to_i32(f32 v) { return __float2int_rz(v); }
WGB_DEV i32 wgb_to_i32(i32 v) { return v; }
WGB_DEV i32 wgb_to_i32(u32 v) { return (i32)v; }
WGB_DEV i32 wgb_to_i32(bool v) { return v ? 1 : 0; }
WGB_DEV u32 wgb_to_u32(f32 v) { return __float2uint_rz(v); }
WGB_DEV u32 wgb_to_u32(i32 v) { return (u32)v; }
WGB_DEV u32 wgb_to_u32(u32 v) { return v; }
WGB_DEV u32 wgb_to_u32(bool v) { return v ? 1u : 0u; }
WGB_DEV bool wgb_to_bool(f32 v) { return v != 0.0f; }
WGB_DEV bool wgb_to_bool(i32 v) { return v != 0; }
WGB_DEV bool wgb_to_bool(u32 v) { return v != 0u; }
WGB_DEV bool wgb_to_bool(bool v) { return v; }
// f16: to it through binary32 (integers up to binary16's largest finite value are exact there, larger ones round to
// infinity either way), from it exactly
WGB_DEV f16 wgb_to_f16(f16 v) { return v; }
WGB_DEV f16 wgb_to_f16(f32 v) { return wgb_f16(v); }
WGB_DEV f16 wgb_to_f16(i32 v) { return wgb_f16(__int2float_rn(v)); }
WGB_DEV f16 wgb_to_f16(u32 v) { return wgb_f16(__uint2float_rn(v)); }
WGB_DEV f16 wgb_to_f16(bool v) { return wgb_f16(v ? 1.0f : 0.0f); }
WGB_DEV f32 wgb_to_f32(f16 v) { return wgb_f32(v); }
WGB_DEV i32 wgb_to_i32(f16 v) { return __float2int_rz(wgb_f32(v)); }
WGB_DEV u32 wgb_to_u32(f16 v) { return __float2uint_rz(wgb_f32(v)); }
WGB_DEV bool wgb_to_bool(f16 v) { return wgb_f32(v) != 0.0f; }

// integer division / remainder: the reference aborts on division by zero and on
// i32::MIN / -1 (binary.rs:451-563); WGSL's defined results are used here instead
WGB_DEV i32 wgb_idiv(i32 a, i32 b) { return (b == 0 || (a == (-2147483647 - 1) && b == -1)) ? a : a / b; }
WGB_DEV u32 wgb_idiv(u32 a, u32 b) { return b == 0u ? a : a / b; }
WGB_DEV i32 wgb_irem(i32 a, i32 b) { return (b == 0 || (a == (-2147483647 - 1) && b == -1)) ? 0 : a % b; }
WGB_DEV u32 wgb_irem(u32 a, u32 b) { return b == 0u ? 0u : a % b; }

// component-wise lifting of scalar functions to vectors (and vector (x) scalar)
#define WGB_LIFT1(FN, S)                                                                            \
    WGB_DEV vec2##S FN(vec2##S a) { return vec2##S(FN(a.x), FN(a.y)); }                               \
    WGB_DEV vec3##S FN(vec3##S a) { return vec3##S(FN(a.x), FN(a.y), FN(a.z)); }                      \
    WGB_DEV vec4##S FN(vec4##S a) { return vec4##S(FN(a.x), FN(a.y), FN(a.z), FN(a.w)); }
#define WGB_LIFT2(FN, S, T)                                                                         \
    WGB_DEV vec2##S FN(vec2##S a, vec2##S b) { return vec2##S(FN(a.x, b.x), FN(a.y, b.y)); }          \
    WGB_DEV vec3##S FN(vec3##S a, vec3##S b) { return vec3##S(FN(a.x, b.x), FN(a.y, b.y), FN(a.z, b.z)); } \
    WGB_DEV vec4##S FN(vec4##S a, vec4##S b) { return vec4##S(FN(a.x, b.x), FN(a.y, b.y), FN(a.z, b.z), FN(a.w, b.w)); } \
    WGB_DEV vec2##S FN(vec2##S a, T b) { return vec2##S(FN(a.x, b), FN(a.y, b)); }                    \
    WGB_DEV vec3##S FN(vec3##S a, T b) { return vec3##S(FN(a.x, b), FN(a.y, b), FN(a.z, b)); }        \
    WGB_DEV vec4##S FN(vec4##S a, T b) { return vec4##S(FN(a.x, b), FN(a.y, b), FN(a.z, b), FN(a.w, b)); } \
    WGB_DEV vec2##S FN(T a, vec2##S b) { return vec2##S(FN(a, b.x), FN(a, b.y)); }                    \
    WGB_DEV vec3##S FN(T a, vec3##S b) { return vec3##S(FN(a, b.x), FN(a, b.y), FN(a, b.z)); }        \
    WGB_DEV vec4##S FN(T a, vec4##S b) { return vec4##S(FN(a, b.x), FN(a, b.y), FN(a, b.z), FN(a, b.w)); }
WGB_LIFT2(wgb_idiv, i, i32)
WGB_LIFT2(wgb_idiv, u, u32)
WGB_LIFT2(wgb_irem, i, i32)
WGB_LIFT2(wgb_irem, u, u32)
WGB_LIFT2(wgb_rem, f, f32)
WGB_LIFT2(wgb_rem, h, f16)
WGB_DEV vec2i operator-(vec2i a) { return vec2i(-a.x, -a.y); }
WGB_DEV vec3i operator-(vec3i a) { return vec3i(-a.x, -a.y, -a.z); }
WGB_DEV vec4i operator-(vec4i a) { return vec4i(-a.x, -a.y, -a.z, -a.w); }

// vector comparisons -> vecNb (expression/binary.rs:172-209 is scalar-only in the reference)
#define WGB_VEC_CMP(NAME, OP, S)                                                                    \
    WGB_DEV vec2b NAME(vec2##S a, vec2##S b) { return vec2b(a.x OP b.x, a.y OP b.y); }                \
    WGB_DEV vec3b NAME(vec3##S a, vec3##S b) { return vec3b(a.x OP b.x, a.y OP b.y, a.z OP b.z); }    \
    WGB_DEV vec4b NAME(vec4##S a, vec4##S b) { return vec4b(a.x OP b.x, a.y OP b.y, a.z OP b.z, a.w OP b.w); }
#define WGB_VEC_CMP_ALL(S) WGB_VEC_CMP(wgb_eq, ==, S) WGB_VEC_CMP(wgb_ne, !=, S) WGB_VEC_CMP(wgb_lt, <, S) \
    WGB_VEC_CMP(wgb_gt, >, S) WGB_VEC_CMP(wgb_le, <=, S) WGB_VEC_CMP(wgb_ge, >=, S)
WGB_VEC_CMP_ALL(f)
WGB_VEC_CMP_ALL(h)
WGB_VEC_CMP_ALL(i)
WGB_VEC_CMP_ALL(u)
WGB_VEC_CMP(wgb_eq, ==, b)
WGB_VEC_CMP(wgb_ne, !=, b)
WGB_DEV vec2b wgb_not(vec2b a) { return vec2b(!a.x, !a.y); }
WGB_DEV vec3b wgb_not(vec3b a) { return vec3b(!a.x, !a.y, !a.z); }
WGB_DEV vec4b wgb_not(vec4b a) { return vec4b(!a.x, !a.y, !a.z, !a.w); }
WGB_DEV bool wgb_all(bool a) { return a; }
WGB_DEV bool wgb_all(vec2b a) { return a.x && a.y; }
WGB_DEV bool wgb_all(vec3b a) { return a.x && a.y && a.z; }
WGB_DEV bool wgb_all(vec4b a) { return a.x && a.y && a.z && a.w; }
WGB_DEV bool wgb_any(bool a) { return a; }
WGB_DEV bool wgb_any(vec2b a) { return a.x || a.y; }
WGB_DEV bool wgb_any(vec3b a) { return a.x || a.y || a.z; }
WGB_DEV bool wgb_any(vec4b a) { return a.x || a.y || a.z || a.w; }
#define WGB_INT_BITOPS(S, T)                                                                        \
    WGB_DEV T wgb_and(T a, T b) { return a & b; } WGB_DEV T wgb_or(T a, T b) { return a | b; } WGB_DEV T wgb_xor(T a, T b) { return a ^ b; } \
    WGB_DEV T wgb_bitnot(T a) { return ~a; }                                                         \
    WGB_DEV T wgb_shl(T a, u32 b) { return a << (b & 31u); } WGB_DEV T wgb_shr(T a, u32 b) { return a >> (b & 31u); } \
    WGB_LIFT2(wgb_and, S, T) WGB_LIFT2(wgb_or, S, T) WGB_LIFT2(wgb_xor, S, T) WGB_LIFT1(wgb_bitnot, S) \
    WGB_DEV vec2##S wgb_shl(vec2##S a, vec2u b) { return vec2##S(wgb_shl(a.x, b.x), wgb_shl(a.y, b.y)); } \
    WGB_DEV vec3##S wgb_shl(vec3##S a, vec3u b) { return vec3##S(wgb_shl(a.x, b.x), wgb_shl(a.y, b.y), wgb_shl(a.z, b.z)); } \
    WGB_DEV vec4##S wgb_shl(vec4##S a, vec4u b) { return vec4##S(wgb_shl(a.x, b.x), wgb_shl(a.y, b.y), wgb_shl(a.z, b.z), wgb_shl(a.w, b.w)); } \
    WGB_DEV vec2##S wgb_shr(vec2##S a, vec2u b) { return vec2##S(wgb_shr(a.x, b.x), wgb_shr(a.y, b.y)); } \
    WGB_DEV vec3##S wgb_shr(vec3##S a, vec3u b) { return vec3##S(wgb_shr(a.x, b.x), wgb_shr(a.y, b.y), wgb_shr(a.z, b.z)); } \
    WGB_DEV vec4##S wgb_shr(vec4##S a, vec4u b) { return vec4##S(wgb_shr(a.x, b.x), wgb_shr(a.y, b.y), wgb_shr(a.z, b.z), wgb_shr(a.w, b.w)); }
WGB_INT_BITOPS(i, i32)
WGB_INT_BITOPS(u, u32)

// select(f, t, cond) (expression/select.rs:36-103: scalar condition; vector conditions select per component)
template <class T> WGB_DEV T wgb_select(T f, T t, bool cond) { return cond ? t : f; }
#define WGB_VEC_SELECT(S)                                                                           \
    WGB_DEV vec2##S wgb_select(vec2##S f, vec2##S t, vec2b c) { return vec2##S(c.x ? t.x : f.x, c.y ? t.y : f.y); } \
    WGB_DEV vec3##S wgb_select(vec3##S f, vec3##S t, vec3b c) { return vec3##S(c.x ? t.x : f.x, c.y ? t.y : f.y, c.z ? t.z : f.z); } \
    WGB_DEV vec4##S wgb_select(vec4##S f, vec4##S t, vec4b c) { return vec4##S(c.x ? t.x : f.x, c.y ? t.y : f.y, c.z ? t.z : f.z, c.w ? t.w : f.w); }
WGB_VEC_SELECT(f)
WGB_VEC_SELECT(h)
WGB_VEC_SELECT(i)
WGB_VEC_SELECT(u)
WGB_VEC_SELECT(b)

template <class To, class From> WGB_DEV To wgb_bitcast(From v) { static_assert(sizeof(To) == sizeof(From), "bitcast size"); To r; memcpy(&r, &v, sizeof(To)); return r; }

// fixed-size arrays by value
template <class T, int N> struct wgb_array {
    T v[N];
    WGB_DEV T& operator[](u32 i) { return v[i]; }
    WGB_DEV const T& operator[](u32 i) const { return v[i]; }
};

// ---------------------------------------------------------------------------------------
// math builtins.  The reference leaves every one of them as todo!() (expression/math.rs:23,29),
// so there is no reference arithmetic to match; they are IEEE single precision, unfused.
// ---------------------------------------------------------------------------------------
WGB_DEV f32 wgb_abs(f32 a) { return fabsf(a); }
WGB_DEV i32 wgb_abs(i32 a) { return a < 0 ? -a : a; }
WGB_DEV u32 wgb_abs(u32 a) { return a; }
WGB_DEV f32 wgb_min(f32 a, f32 b) { return fminf(a, b); }
WGB_DEV i32 wgb_min(i32 a, i32 b) { return a < b ? a : b; }
WGB_DEV u32 wgb_min(u32 a, u32 b) { return a < b ? a : b; }
WGB_DEV f32 wgb_max(f32 a, f32 b) { return fmaxf(a, b); }
WGB_DEV i32 wgb_max(i32 a, i32 b) { return a > b ? a : b; }
WGB_DEV u32 wgb_max(u32 a, u32 b) { return a > b ? a : b; }
WGB_DEV f32 wgb_floor(f32 a) { return floorf(a); }
WGB_DEV f32 wgb_ceil(f32 a) { return ceilf(a); }
WGB_DEV f32 wgb_round(f32 a) { return rintf(a); }          // WGSL round: ties to even
WGB_DEV f32 wgb_trunc(f32 a) { return truncf(a); }
WGB_DEV f32 wgb_fract(f32 a) { return __fsub_rn(a, floorf(a)); }
WGB_DEV f32 wgb_sqrt(f32 a) { return __fsqrt_rn(a); }
WGB_DEV f32 wgb_inverseSqrt(f32 a) { return __fdiv_rn(1.0f, __fsqrt_rn(a)); }
WGB_DEV f32 wgb_sin(f32 a) { return sinf(a); }
WGB_DEV f32 wgb_cos(f32 a) { return cosf(a); }
WGB_DEV f32 wgb_tan(f32 a) { return tanf(a); }
WGB_DEV f32 wgb_asin(f32 a) { return asinf(a); }
WGB_DEV f32 wgb_acos(f32 a) { return acosf(a); }
WGB_DEV f32 wgb_atan(f32 a) { return atanf(a); }
WGB_DEV f32 wgb_atan2(f32 a, f32 b) { return atan2f(a, b); }
WGB_DEV f32 wgb_sinh(f32 a) { return sinhf(a); }
WGB_DEV f32 wgb_cosh(f32 a) { return coshf(a); }
WGB_DEV f32 wgb_tanh(f32 a) { return tanhf(a); }
WGB_DEV f32 wgb_exp(f32 a) { return expf(a); }
WGB_DEV f32 wgb_exp2(f32 a) { return exp2f(a); }
WGB_DEV f32 wgb_log(f32 a) { return logf(a); }
WGB_DEV f32 wgb_log2(f32 a) { return log2f(a); }
WGB_DEV f32 wgb_pow(f32 a, f32 b) { return powf(a, b); }
WGB_DEV f32 wgb_sign(f32 a) { return a > 0.0f ? 1.0f : (a < 0.0f ? -1.0f : 0.0f); }
WGB_DEV i32 wgb_sign(i32 a) { return a > 0 ? 1 : (a < 0 ? -1 : 0); }
WGB_DEV f32 wgb_step(f32 edge, f32 x) { return x >= edge ? 1.0f : 0.0f; }
WGB_DEV f32 wgb_saturate(f32 a) { return fminf(fmaxf(a, 0.0f), 1.0f); }
WGB_DEV f32 wgb_degrees(f32 a) { return __fmul_rn(a, 57.295779513082322865f); }
WGB_DEV f32 wgb_radians(f32 a) { return __fmul_rn(a, 0.017453292519943295474f); }
WGB_DEV f32 wgb_clamp(f32 x, f32 lo, f32 hi) { return fminf(fmaxf(x, lo), hi); }
WGB_DEV i32 wgb_clamp(i32 x, i32 lo, i32 hi) { return wgb_min(wgb_max(x, lo), hi); }
WGB_DEV u32 wgb_clamp(u32 x, u32 lo, u32 hi) { return wgb_min(wgb_max(x, lo), hi); }
WGB_DEV f32 wgb_mix(f32 a, f32 b, f32 t) { return __fadd_rn(__fmul_rn(a, __fsub_rn(1.0f, t)), __fmul_rn(b, t)); }
WGB_DEV f32 wgb_fma(f32 a, f32 b, f32 c) { return __fmaf_rn(a, b, c); }
WGB_DEV f32 wgb_smoothstep(f32 lo, f32 hi, f32 x) {
    const f32 t = wgb_saturate(__fdiv_rn(__fsub_rn(x, lo), __fsub_rn(hi, lo)));
    return __fmul_rn(__fmul_rn(t, t), __fsub_rn(3.0f, __fmul_rn(2.0f, t)));
}
#define WGB_LIFT1F(FN) WGB_LIFT1(FN, f)
WGB_LIFT1F(wgb_abs) WGB_LIFT1F(wgb_floor) WGB_LIFT1F(wgb_ceil) WGB_LIFT1F(wgb_round) WGB_LIFT1F(wgb_trunc) WGB_LIFT1F(wgb_fract)
WGB_LIFT1F(wgb_sqrt) WGB_LIFT1F(wgb_inverseSqrt) WGB_LIFT1F(wgb_sin) WGB_LIFT1F(wgb_cos) WGB_LIFT1F(wgb_tan) WGB_LIFT1F(wgb_asin)
WGB_LIFT1F(wgb_acos) WGB_LIFT1F(wgb_atan) WGB_LIFT1F(wgb_sinh) WGB_LIFT1F(wgb_cosh) WGB_LIFT1F(wgb_tanh) WGB_LIFT1F(wgb_exp)
WGB_LIFT1F(wgb_exp2) WGB_LIFT1F(wgb_log) WGB_LIFT1F(wgb_log2) WGB_LIFT1F(wgb_sign) WGB_LIFT1F(wgb_saturate) WGB_LIFT1F(wgb_degrees)
WGB_LIFT1F(wgb_radians)
WGB_LIFT1(wgb_abs, i) WGB_LIFT1(wgb_sign, i)
WGB_LIFT2(wgb_min, f, f32) WGB_LIFT2(wgb_max, f, f32) WGB_LIFT2(wgb_atan2, f, f32) WGB_LIFT2(wgb_pow, f, f32) WGB_LIFT2(wgb_step, f, f32)
WGB_LIFT2(wgb_min, i, i32) WGB_LIFT2(wgb_max, i, i32) WGB_LIFT2(wgb_min, u, u32) WGB_LIFT2(wgb_max, u, u32)
#define WGB_LIFT3(FN, S, T)                                                                         \
    WGB_DEV vec2##S FN(vec2##S a, vec2##S b, vec2##S c) { return vec2##S(FN(a.x, b.x, c.x), FN(a.y, b.y, c.y)); } \
    WGB_DEV vec3##S FN(vec3##S a, vec3##S b, vec3##S c) { return vec3##S(FN(a.x, b.x, c.x), FN(a.y, b.y, c.y), FN(a.z, b.z, c.z)); } \
    WGB_DEV vec4##S FN(vec4##S a, vec4##S b, vec4##S c) { return vec4##S(FN(a.x, b.x, c.x), FN(a.y, b.y, c.y), FN(a.z, b.z, c.z), FN(a.w, b.w, c.w)); } \
    WGB_DEV vec2##S FN(vec2##S a, T b, T c) { return vec2##S(FN(a.x, b, c), FN(a.y, b, c)); }        \
    WGB_DEV vec3##S FN(vec3##S a, T b, T c) { return vec3##S(FN(a.x, b, c), FN(a.y, b, c), FN(a.z, b, c)); } \
    WGB_DEV vec4##S FN(vec4##S a, T b, T c) { return vec4##S(FN(a.x, b, c), FN(a.y, b, c), FN(a.z, b, c), FN(a.w, b, c)); }
WGB_LIFT3(wgb_clamp, f, f32) WGB_LIFT3(wgb_clamp, i, i32) WGB_LIFT3(wgb_clamp, u, u32) WGB_LIFT3(wgb_fma, f, f32)
// f16: the selections are exact (the result is one of the operands); the other math builtins are not offered for f16
WGB_DEV f16 wgb_abs(f16 a) { f16 r; r.bits = (unsigned short)(a.bits & 0x7FFFu); return r; }
WGB_DEV f16 wgb_min(f16 a, f16 b) { return wgb_f16(fminf(wgb_f32(a), wgb_f32(b))); }
WGB_DEV f16 wgb_max(f16 a, f16 b) { return wgb_f16(fmaxf(wgb_f32(a), wgb_f32(b))); }
WGB_DEV f16 wgb_clamp(f16 x, f16 lo, f16 hi) { return wgb_min(wgb_max(x, lo), hi); }
WGB_LIFT1(wgb_abs, h) WGB_LIFT2(wgb_min, h, f16) WGB_LIFT2(wgb_max, h, f16) WGB_LIFT3(wgb_clamp, h, f16)
// ---- integer bit builtins (WGSL 17.5.x; all `todo!()` in the reference with the rest of Math) ----
WGB_DEV u32 wgb_countOneBits(u32 a) { return (u32)__popc(a); }
WGB_DEV i32 wgb_countOneBits(i32 a) { return __popc((u32)a); }
WGB_DEV u32 wgb_countLeadingZeros(u32 a) { return (u32)__clz((int)a); }
WGB_DEV i32 wgb_countLeadingZeros(i32 a) { return __clz(a); }
WGB_DEV u32 wgb_countTrailingZeros(u32 a) { return a == 0u ? 32u : (u32)(__ffs((int)a) - 1); }
WGB_DEV i32 wgb_countTrailingZeros(i32 a) { return a == 0 ? 32 : __ffs(a) - 1; }
WGB_DEV u32 wgb_firstLeadingBit(u32 a) { return a == 0u ? 0xFFFFFFFFu : 31u - (u32)__clz((int)a); }
WGB_DEV i32 wgb_firstLeadingBit(i32 a) { return (a == 0 || a == -1) ? -1 : 31 - __clz(a < 0 ? ~a : a); }
WGB_DEV u32 wgb_firstTrailingBit(u32 a) { return a == 0u ? 0xFFFFFFFFu : (u32)(__ffs((int)a) - 1); }
WGB_DEV i32 wgb_firstTrailingBit(i32 a) { return a == 0 ? -1 : __ffs(a) - 1; }
WGB_DEV u32 wgb_reverseBits(u32 a) { return __brev(a); }
WGB_DEV i32 wgb_reverseBits(i32 a) { return (i32)__brev((u32)a); }
WGB_DEV u32 wgb_extractBits(u32 e, u32 offset, u32 count) {
    const u32 o = offset < 32u ? offset : 32u, c = count < 32u - o ? count : 32u - o;
    return c == 0u ? 0u : (e >> o) & (c == 32u ? 0xFFFFFFFFu : (1u << c) - 1u);
}
WGB_DEV i32 wgb_extractBits(i32 e, u32 offset, u32 count) {
    const u32 o = offset < 32u ? offset : 32u, c = count < 32u - o ? count : 32u - o;
    return c == 0u ? 0 : ((i32)((u32)e << (32u - c - o))) >> (32u - c);          // sign-extends
}
WGB_DEV u32 wgb_insertBits(u32 e, u32 newbits, u32 offset, u32 count) {
    const u32 o = offset < 32u ? offset : 32u, c = count < 32u - o ? count : 32u - o;
    if (c == 0u) return e;
    const u32 mask = (c == 32u ? 0xFFFFFFFFu : (1u << c) - 1u) << o;
    return (e & ~mask) | ((newbits << o) & mask);
}
WGB_DEV i32 wgb_insertBits(i32 e, i32 newbits, u32 offset, u32 count) { return (i32)wgb_insertBits((u32)e, (u32)newbits, offset, count); }
WGB_LIFT1(wgb_countOneBits, i) WGB_LIFT1(wgb_countOneBits, u) WGB_LIFT1(wgb_countLeadingZeros, i) WGB_LIFT1(wgb_countLeadingZeros, u)
WGB_LIFT1(wgb_countTrailingZeros, i) WGB_LIFT1(wgb_countTrailingZeros, u) WGB_LIFT1(wgb_firstLeadingBit, i) WGB_LIFT1(wgb_firstLeadingBit, u)
WGB_LIFT1(wgb_firstTrailingBit, i) WGB_LIFT1(wgb_firstTrailingBit, u) WGB_LIFT1(wgb_reverseBits, i) WGB_LIFT1(wgb_reverseBits, u)
#define WGB_LIFT_BITS(S)                                                                            \
    WGB_DEV vec2##S wgb_extractBits(vec2##S e, u32 o, u32 c) { return vec2##S(wgb_extractBits(e.x, o, c), wgb_extractBits(e.y, o, c)); } \
    WGB_DEV vec3##S wgb_extractBits(vec3##S e, u32 o, u32 c) { return vec3##S(wgb_extractBits(e.x, o, c), wgb_extractBits(e.y, o, c), wgb_extractBits(e.z, o, c)); } \
    WGB_DEV vec4##S wgb_extractBits(vec4##S e, u32 o, u32 c) { return vec4##S(wgb_extractBits(e.x, o, c), wgb_extractBits(e.y, o, c), wgb_extractBits(e.z, o, c), wgb_extractBits(e.w, o, c)); } \
    WGB_DEV vec2##S wgb_insertBits(vec2##S e, vec2##S n, u32 o, u32 c) { return vec2##S(wgb_insertBits(e.x, n.x, o, c), wgb_insertBits(e.y, n.y, o, c)); } \
    WGB_DEV vec3##S wgb_insertBits(vec3##S e, vec3##S n, u32 o, u32 c) { return vec3##S(wgb_insertBits(e.x, n.x, o, c), wgb_insertBits(e.y, n.y, o, c), wgb_insertBits(e.z, n.z, o, c)); } \
    WGB_DEV vec4##S wgb_insertBits(vec4##S e, vec4##S n, u32 o, u32 c) { return vec4##S(wgb_insertBits(e.x, n.x, o, c), wgb_insertBits(e.y, n.y, o, c), wgb_insertBits(e.z, n.z, o, c), wgb_insertBits(e.w, n.w, o, c)); }
WGB_LIFT_BITS(i) WGB_LIFT_BITS(u)
WGB_DEV vec2f wgb_mix(vec2f a, vec2f b, vec2f t) { return vec2f(wgb_mix(a.x, b.x, t.x), wgb_mix(a.y, b.y, t.y)); }
WGB_DEV vec3f wgb_mix(vec3f a, vec3f b, vec3f t) { return vec3f(wgb_mix(a.x, b.x, t.x), wgb_mix(a.y, b.y, t.y), wgb_mix(a.z, b.z, t.z)); }
WGB_DEV vec4f wgb_mix(vec4f a, vec4f b, vec4f t) { return vec4f(wgb_mix(a.x, b.x, t.x), wgb_mix(a.y, b.y, t.y), wgb_mix(a.z, b.z, t.z), wgb_mix(a.w, b.w, t.w)); }
WGB_DEV vec2f wgb_mix(vec2f a, vec2f b, f32 t) { return wgb_mix(a, b, vec2f(t)); }
WGB_DEV vec3f wgb_mix(vec3f a, vec3f b, f32 t) { return wgb_mix(a, b, vec3f(t)); }
WGB_DEV vec4f wgb_mix(vec4f a, vec4f b, f32 t) { return wgb_mix(a, b, vec4f(t)); }
WGB_DEV vec2f wgb_smoothstep(vec2f a, vec2f b, vec2f x) { return vec2f(wgb_smoothstep(a.x, b.x, x.x), wgb_smoothstep(a.y, b.y, x.y)); }
WGB_DEV vec3f wgb_smoothstep(vec3f a, vec3f b, vec3f x) { return vec3f(wgb_smoothstep(a.x, b.x, x.x), wgb_smoothstep(a.y, b.y, x.y), wgb_smoothstep(a.z, b.z, x.z)); }
WGB_DEV vec4f wgb_smoothstep(vec4f a, vec4f b, vec4f x) { return vec4f(wgb_smoothstep(a.x, b.x, x.x), wgb_smoothstep(a.y, b.y, x.y), wgb_smoothstep(a.z, b.z, x.z), wgb_smoothstep(a.w, b.w, x.w)); }
WGB_DEV vec2f wgb_smoothstep(f32 a, f32 b, vec2f x) { return wgb_smoothstep(vec2f(a), vec2f(b), x); }
WGB_DEV vec3f wgb_smoothstep(f32 a, f32 b, vec3f x) { return wgb_smoothstep(vec3f(a), vec3f(b), x); }
WGB_DEV vec4f wgb_smoothstep(f32 a, f32 b, vec4f x) { return wgb_smoothstep(vec4f(a), vec4f(b), x); }
WGB_DEV f32 wgb_dot(vec2f a, vec2f b) { return __fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)); }
WGB_DEV f32 wgb_dot(vec3f a, vec3f b) { return __fadd_rn(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z)); }
WGB_DEV f32 wgb_dot(vec4f a, vec4f b) { return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z)), __fmul_rn(a.w, b.w)); }
WGB_DEV i32 wgb_dot(vec2i a, vec2i b) { return a.x * b.x + a.y * b.y; }
WGB_DEV i32 wgb_dot(vec3i a, vec3i b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
WGB_DEV i32 wgb_dot(vec4i a, vec4i b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
WGB_DEV u32 wgb_dot(vec2u a, vec2u b) { return a.x * b.x + a.y * b.y; }
WGB_DEV u32 wgb_dot(vec3u a, vec3u b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
WGB_DEV u32 wgb_dot(vec4u a, vec4u b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
WGB_DEV vec3f wgb_cross(vec3f a, vec3f b) {
    return vec3f(__fsub_rn(__fmul_rn(a.y, b.z), __fmul_rn(a.z, b.y)), __fsub_rn(__fmul_rn(a.z, b.x), __fmul_rn(a.x, b.z)),
                 __fsub_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
}
WGB_DEV f32 wgb_length(f32 a) { return fabsf(a); }
WGB_DEV f32 wgb_length(vec2f a) { return __fsqrt_rn(wgb_dot(a, a)); }
WGB_DEV f32 wgb_length(vec3f a) { return __fsqrt_rn(wgb_dot(a, a)); }
WGB_DEV f32 wgb_length(vec4f a) { return __fsqrt_rn(wgb_dot(a, a)); }
WGB_DEV f32 wgb_distance(f32 a, f32 b) { return fabsf(__fsub_rn(a, b)); }
WGB_DEV f32 wgb_distance(vec2f a, vec2f b) { return wgb_length(a - b); }
WGB_DEV f32 wgb_distance(vec3f a, vec3f b) { return wgb_length(a - b); }
WGB_DEV f32 wgb_distance(vec4f a, vec4f b) { return wgb_length(a - b); }
WGB_DEV vec2f wgb_normalize(vec2f a) { return a / wgb_length(a); }
WGB_DEV vec3f wgb_normalize(vec3f a) { return a / wgb_length(a); }
WGB_DEV vec4f wgb_normalize(vec4f a) { return a / wgb_length(a); }
WGB_DEV vec2f wgb_reflect(vec2f e, vec2f n) { return e - n * __fmul_rn(2.0f, wgb_dot(n, e)); }
WGB_DEV vec3f wgb_reflect(vec3f e, vec3f n) { return e - n * __fmul_rn(2.0f, wgb_dot(n, e)); }
WGB_DEV vec4f wgb_reflect(vec4f e, vec4f n) { return e - n * __fmul_rn(2.0f, wgb_dot(n, e)); }

// ---------------------------------------------------------------------------------------
// resources
// ---------------------------------------------------------------------------------------
// arrayLength(&binding.array): elements that fit between the array's offset and the end of the bound range
WGB_DEV u32 wgb_array_length(const WgbDraw& d, int g, int b, u32 offset, u32 stride) {
    const u32 size = d.res[g][b].a;
    return size > offset ? (size - offset) / stride : 0u;
}
// uniform / storage loads: `offset` is the WGSL-layout byte offset computed by the emitter
template <class T> WGB_DEV T wgb_load(const WgbDraw& d, int group, int binding, u32 offset);
template <> WGB_DEV f32 wgb_load<f32>(const WgbDraw& d, int g, int b, u32 off) {
    return __ldg(reinterpret_cast<const f32*>(d.res[g][b].ptr + off));
}
template <> WGB_DEV i32 wgb_load<i32>(const WgbDraw& d, int g, int b, u32 off) {
    return __ldg(reinterpret_cast<const i32*>(d.res[g][b].ptr + off));
}
template <> WGB_DEV u32 wgb_load<u32>(const WgbDraw& d, int g, int b, u32 off) {
    return __ldg(reinterpret_cast<const u32*>(d.res[g][b].ptr + off));
}
template <> WGB_DEV vec2f wgb_load<vec2f>(const WgbDraw& d, int g, int b, u32 off) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(d.res[g][b].ptr + off));
    return vec2f(v.x, v.y);
}
template <> WGB_DEV vec3f wgb_load<vec3f>(const WgbDraw& d, int g, int b, u32 off) {
    const f32* p = reinterpret_cast<const f32*>(d.res[g][b].ptr + off);
    return vec3f(__ldg(p), __ldg(p + 1), __ldg(p + 2));
}
template <> WGB_DEV vec4f wgb_load<vec4f>(const WgbDraw& d, int g, int b, u32 off) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(d.res[g][b].ptr + off));
    return vec4f(v.x, v.y, v.z, v.w);
}
template <> WGB_DEV vec2i wgb_load<vec2i>(const WgbDraw& d, int g, int b, u32 off) { return vec2i(wgb_load<i32>(d, g, b, off), wgb_load<i32>(d, g, b, off + 4)); }
template <> WGB_DEV vec3i wgb_load<vec3i>(const WgbDraw& d, int g, int b, u32 off) { return vec3i(wgb_load<i32>(d, g, b, off), wgb_load<i32>(d, g, b, off + 4), wgb_load<i32>(d, g, b, off + 8)); }
template <> WGB_DEV vec4i wgb_load<vec4i>(const WgbDraw& d, int g, int b, u32 off) { return vec4i(wgb_load<i32>(d, g, b, off), wgb_load<i32>(d, g, b, off + 4), wgb_load<i32>(d, g, b, off + 8), wgb_load<i32>(d, g, b, off + 12)); }
template <> WGB_DEV vec2u wgb_load<vec2u>(const WgbDraw& d, int g, int b, u32 off) { return vec2u(wgb_load<u32>(d, g, b, off), wgb_load<u32>(d, g, b, off + 4)); }
template <> WGB_DEV vec3u wgb_load<vec3u>(const WgbDraw& d, int g, int b, u32 off) { return vec3u(wgb_load<u32>(d, g, b, off), wgb_load<u32>(d, g, b, off + 4), wgb_load<u32>(d, g, b, off + 8)); }
template <> WGB_DEV vec4u wgb_load<vec4u>(const WgbDraw& d, int g, int b, u32 off) { return vec4u(wgb_load<u32>(d, g, b, off), wgb_load<u32>(d, g, b, off + 4), wgb_load<u32>(d, g, b, off + 8), wgb_load<u32>(d, g, b, off + 12)); }
// matrix columns are 16-byte aligned for 3 and 4 rows, 8-byte aligned for 2 rows (WGSL layout)
template <> WGB_DEV mat3x3f wgb_load<mat3x3f>(const WgbDraw& d, int g, int b, u32 off) {
    return mat3x3f(wgb_load<vec3f>(d, g, b, off), wgb_load<vec3f>(d, g, b, off + 16), wgb_load<vec3f>(d, g, b, off + 32));
}
template <> WGB_DEV mat2x2f wgb_load<mat2x2f>(const WgbDraw& d, int g, int b, u32 off) {
    return mat2x2f(wgb_load<vec2f>(d, g, b, off), wgb_load<vec2f>(d, g, b, off + 8));
}
template <> WGB_DEV mat4x4f wgb_load<mat4x4f>(const WgbDraw& d, int g, int b, u32 off) {
    return mat4x4f(wgb_load<vec4f>(d, g, b, off), wgb_load<vec4f>(d, g, b, off + 16),
                   wgb_load<vec4f>(d, g, b, off + 32), wgb_load<vec4f>(d, g, b, off + 48));
}

// texture formats / address modes (numeric values of include/wgpu_b200.h)
#define WGB_FMT_RGBA8_UNORM 0
#define WGB_FMT_RGBA8_UNORM_SRGB 1
#define WGB_FMT_BGRA8_UNORM 2
#define WGB_FMT_BGRA8_UNORM_SRGB 3
#define WGB_FMT_R8_UNORM 4
#define WGB_FMT_RG8_UNORM 5
#define WGB_FMT_RGBA8_SNORM 6
#define WGB_FMT_DEPTH32_FLOAT 7
#define WGB_ADDR_CLAMP_TO_EDGE 0
#define WGB_ADDR_REPEAT 1
#define WGB_ADDR_MIRROR_REPEAT 2

// f32::rem_euclid (Rust std): r = fmod(x, rhs); r < 0 ? r + |rhs| : r
WGB_DEV f32 wgb_rem_euclid(f32 x, f32 rhs) {
    f32 r;
    if (rhs == 1.0f) {
        // fmod(x, 1) = x - trunc(x): both sides are exact (the difference of x and its integer part is representable),
        // a zero result takes the sign of x as fmod's does, infinities give NaN either way
        r = __fsub_rn(x, truncf(x));
        if (r == 0.0f) r = copysignf(0.0f, x);
    } else r = fmodf(x, rhs);
    return r < 0.0f ? __fadd_rn(r, fabsf(rhs)) : r;
}
// Rust `as u32` on f32: saturating, NaN -> 0 (== cvt.rzi.u32.f32)
WGB_DEV u32 wgb_f32_as_u32(f32 v) { return __float2uint_rz(v); }

// texel_coordinate (wgpu-cpu/src/render_pass/binding.rs:151-164)
WGB_DEV u32 wgb_texel_coordinate(f32 x, u32 address_mode, u32 size) {
    if (address_mode == WGB_ADDR_CLAMP_TO_EDGE) {
        if (x < 0.0f) x = 0.0f;          // f32::clamp keeps NaN
        if (x > 1.0f) x = 1.0f;
    } else if (address_mode == WGB_ADDR_REPEAT) {
        x = wgb_rem_euclid(x, 1.0f);
    } else {
        const f32 r = wgb_rem_euclid(x, 2.0f);
        x = (r <= 1.0f) ? r : __fsub_rn(2.0f, r);
    }
    // f32::round: half away from zero
    return wgb_f32_as_u32(roundf(__fmul_rn(x, __uint2float_rn(size - 1u))));
}

// u8 as f32 / 255.0 (texture.rs:170-188), correctly rounded: the quotient from the rounded reciprocal of 255 and one
// residual correction -- the fast path a compiler emits for div.rn.f32, without its range check, which 0..255 over 255
// cannot fail (tests/test_parity_gpu.py::test_every_texel_value_decodes_as_the_division_does)
WGB_DEV f32 wgb_unorm8(u32 v) {
    const f32 a = __uint2float_rn(v), r = 0.003921568859368563f;      // RN(1 / 255) = 0x3B808081
    const f32 q = __fmaf_rn(a, r, 0.0f);
    return __fmaf_rn(__fmaf_rn(-255.0f, q, a), r, q);
}
// textureSample: nearest, mip 0, 2-D (binding.rs:93-149), texel decode u8 as f32 / 255.0
// without sRGB decode (texture.rs:170-188).  The texel is fetched through the bindless
// texture object by integer element index, so the hardware never rounds or filters.
WGB_DEV vec4f wgb_texture_sample(const WgbDraw& d, int tg, int tb, int sg, int sb, vec2f uv) {
    const WgbResource& t = d.res[tg][tb];
    const WgbResource& s = d.res[sg][sb];
    const u32 tx = wgb_texel_coordinate(uv.x, s.a, t.a);
    const u32 ty = wgb_texel_coordinate(uv.y, s.b, t.b);
    const uchar4 p = tex1Dfetch<uchar4>((cudaTextureObject_t)t.tex, (int)(ty * t.a + tx));
    return vec4f(wgb_unorm8(p.x), wgb_unorm8(p.y), wgb_unorm8(p.z), wgb_unorm8(p.w));
}

WGB_DEV vec2u wgb_texture_dimensions(const WgbDraw& d, int tg, int tb) { return vec2u(d.res[tg][tb].a, d.res[tg][tb].b); }
// textureLoad (ImageLoad is a todo!() in the reference, expression/image.rs:107): out-of-range texels read as zero
template <class V> WGB_DEV vec4f wgb_texture_load(const WgbDraw& d, int tg, int tb, V c) {
    const WgbResource& t = d.res[tg][tb];
    const u32 x = (u32)c.x, y = (u32)c.y;
    if (x >= t.a || y >= t.b) return vec4f();
    const uchar4 p = tex1Dfetch<uchar4>((cudaTextureObject_t)t.tex, (int)(y * t.a + x));
    return vec4f(__fdiv_rn((f32)p.x, 255.0f), __fdiv_rn((f32)p.y, 255.0f), __fdiv_rn((f32)p.z, 255.0f), __fdiv_rn((f32)p.w, 255.0f));
}

// ---------------------------------------------------------------------------------------
// stage I/O glue types
// ---------------------------------------------------------------------------------------
struct WgbFragIn {
    vec4f position;          // (vp.x, vp.y, ndc.z, 1/w)  raster.rs:155
    bool front_facing;
    u32 primitive_index;
    u32 sample_index;
    u32 sample_mask;
};
struct WgbFragOut {
    vec4f color[WGB_MAX_COLOR];
    f32 frag_depth;
};

// vertex attribute fetch; the WGB_ATTR<location>_* macros are generated by the host from
// the pipeline's vertex buffer layouts (vertex.rs:60-85, 128-156, 195-199)
template <class T> WGB_DEV T wgb_fetch_raw(const WgbDraw& d, int slot, u32 stride, u32 offset, bool per_instance,
                                           u32 vertex_index, u32 instance_index, u32& oob);
#define WGB_FETCH_ADDR()                                                                            \
    const wgb_u64 start = (wgb_u64)(per_instance ? instance_index : vertex_index) * stride + offset; \
    if (start + sizeof(T) > d.vb[slot].size) { oob = 1u; return T(); }                              \
    const wgb_u64 addr = d.vb[slot].ptr + start;
template <> WGB_DEV vec4f wgb_fetch_raw<vec4f>(const WgbDraw& d, int slot, u32 stride, u32 offset, bool per_instance,
                                              u32 vertex_index, u32 instance_index, u32& oob) {
    typedef vec4f T;
    WGB_FETCH_ADDR()
    if ((addr & 15u) == 0) {       // 128-bit vectorised fetch whenever the layout allows it
        const float4 v = __ldg(reinterpret_cast<const float4*>(addr));
        return vec4f(v.x, v.y, v.z, v.w);
    }
    if ((addr & 7u) == 0) {
        const float2 a = __ldg(reinterpret_cast<const float2*>(addr));
        const float2 b = __ldg(reinterpret_cast<const float2*>(addr + 8));
        return vec4f(a.x, a.y, b.x, b.y);
    }
    const f32* p = reinterpret_cast<const f32*>(addr);
    return vec4f(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
}
template <> WGB_DEV vec3f wgb_fetch_raw<vec3f>(const WgbDraw& d, int slot, u32 stride, u32 offset, bool per_instance,
                                              u32 vertex_index, u32 instance_index, u32& oob) {
    typedef vec3f T;
    WGB_FETCH_ADDR()
    const f32* p = reinterpret_cast<const f32*>(addr);
    return vec3f(__ldg(p), __ldg(p + 1), __ldg(p + 2));
}
template <> WGB_DEV vec2f wgb_fetch_raw<vec2f>(const WgbDraw& d, int slot, u32 stride, u32 offset, bool per_instance,
                                              u32 vertex_index, u32 instance_index, u32& oob) {
    typedef vec2f T;
    WGB_FETCH_ADDR()
    if ((addr & 7u) == 0) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(addr));
        return vec2f(v.x, v.y);
    }
    const f32* p = reinterpret_cast<const f32*>(addr);
    return vec2f(__ldg(p), __ldg(p + 1));
}
template <> WGB_DEV f32 wgb_fetch_raw<f32>(const WgbDraw& d, int slot, u32 stride, u32 offset, bool per_instance,
                                          u32 vertex_index, u32 instance_index, u32& oob) {
    typedef f32 T;
    WGB_FETCH_ADDR()
    return __ldg(reinterpret_cast<const f32*>(addr));
}
template <> WGB_DEV u32 wgb_fetch_raw<u32>(const WgbDraw& d, int slot, u32 stride, u32 offset, bool per_instance,
                                          u32 vertex_index, u32 instance_index, u32& oob) {
    typedef u32 T;
    WGB_FETCH_ADDR()
    return __ldg(reinterpret_cast<const u32*>(addr));
}
template <> WGB_DEV i32 wgb_fetch_raw<i32>(const WgbDraw& d, int slot, u32 stride, u32 offset, bool per_instance,
                                          u32 vertex_index, u32 instance_index, u32& oob) {
    typedef i32 T;
    WGB_FETCH_ADDR()
    return __ldg(reinterpret_cast<const i32*>(addr));
}
#define WGB_FETCH_VEC(VT, ST, N)                                                                    \
    template <> WGB_DEV VT wgb_fetch_raw<VT>(const WgbDraw& d, int slot, u32 stride, u32 offset, bool per_instance, \
                                             u32 vertex_index, u32 instance_index, u32& oob) {        \
        typedef VT T;                                                                               \
        WGB_FETCH_ADDR()                                                                            \
        const ST* p = reinterpret_cast<const ST*>(addr);                                            \
        VT r;                                                                                       \
        for (int i = 0; i < N; i++) r[i] = __ldg(p + i);                                            \
        return r;                                                                                   \
    }
WGB_FETCH_VEC(vec2u, u32, 2) WGB_FETCH_VEC(vec3u, u32, 3) WGB_FETCH_VEC(vec4u, u32, 4)
WGB_FETCH_VEC(vec2i, i32, 2) WGB_FETCH_VEC(vec3i, i32, 3) WGB_FETCH_VEC(vec4i, i32, 4)
// WGB_FETCH(T, LOC): expands to the fetch for @location(LOC) of the bound pipeline layout
#define WGB_FETCH(T, LOC)                                                                           \
    wgb_fetch_raw<T>(wgb, WGB_ATTR##LOC##_SLOT, WGB_ATTR##LOC##_STRIDE, WGB_ATTR##LOC##_OFFSET,        \
                     WGB_ATTR##LOC##_INSTANCE, vertex_index, instance_index, oob)

// varyings travel between the stages as 32-bit slots (inter-stage layout of
// naga-cranelift/src/bindings.rs:307-346 in units of 4 bytes)
WGB_DEV u32 wgb_bits(f32 v) { return __float_as_uint(v); }
WGB_DEV u32 wgb_bits(u32 v) { return v; }
WGB_DEV u32 wgb_bits(i32 v) { return (u32)v; }
WGB_DEV void wgb_put(u32* s, int o, f32 v) { s[o] = wgb_bits(v); }
WGB_DEV void wgb_put(u32* s, int o, u32 v) { s[o] = v; }
WGB_DEV void wgb_put(u32* s, int o, i32 v) { s[o] = (u32)v; }
WGB_DEV void wgb_put(u32* s, int o, vec2f v) { s[o] = wgb_bits(v.x); s[o + 1] = wgb_bits(v.y); }
WGB_DEV void wgb_put(u32* s, int o, vec3f v) { s[o] = wgb_bits(v.x); s[o + 1] = wgb_bits(v.y); s[o + 2] = wgb_bits(v.z); }
WGB_DEV void wgb_put(u32* s, int o, vec4f v) { s[o] = wgb_bits(v.x); s[o + 1] = wgb_bits(v.y); s[o + 2] = wgb_bits(v.z); s[o + 3] = wgb_bits(v.w); }
WGB_DEV void wgb_put(u32* s, int o, vec2u v) { s[o] = v.x; s[o + 1] = v.y; }
WGB_DEV void wgb_put(u32* s, int o, vec3u v) { s[o] = v.x; s[o + 1] = v.y; s[o + 2] = v.z; }
WGB_DEV void wgb_put(u32* s, int o, vec4u v) { s[o] = v.x; s[o + 1] = v.y; s[o + 2] = v.z; s[o + 3] = v.w; }
WGB_DEV void wgb_put(u32* s, int o, vec2i v) { s[o] = (u32)v.x; s[o + 1] = (u32)v.y; }
WGB_DEV void wgb_put(u32* s, int o, vec3i v) { s[o] = (u32)v.x; s[o + 1] = (u32)v.y; s[o + 2] = (u32)v.z; }
WGB_DEV void wgb_put(u32* s, int o, vec4i v) { s[o] = (u32)v.x; s[o + 1] = (u32)v.y; s[o + 2] = (u32)v.z; s[o + 3] = (u32)v.w; }
template <class T> WGB_DEV T wgb_get(const u32* s, int o);
template <> WGB_DEV f32 wgb_get<f32>(const u32* s, int o) { return __uint_as_float(s[o]); }
template <> WGB_DEV u32 wgb_get<u32>(const u32* s, int o) { return s[o]; }
template <> WGB_DEV i32 wgb_get<i32>(const u32* s, int o) { return (i32)s[o]; }
template <> WGB_DEV vec2f wgb_get<vec2f>(const u32* s, int o) { return vec2f(__uint_as_float(s[o]), __uint_as_float(s[o + 1])); }
template <> WGB_DEV vec3f wgb_get<vec3f>(const u32* s, int o) { return vec3f(__uint_as_float(s[o]), __uint_as_float(s[o + 1]), __uint_as_float(s[o + 2])); }
template <> WGB_DEV vec4f wgb_get<vec4f>(const u32* s, int o) { return vec4f(__uint_as_float(s[o]), __uint_as_float(s[o + 1]), __uint_as_float(s[o + 2]), __uint_as_float(s[o + 3])); }
template <> WGB_DEV vec2u wgb_get<vec2u>(const u32* s, int o) { return vec2u(s[o], s[o + 1]); }
template <> WGB_DEV vec3u wgb_get<vec3u>(const u32* s, int o) { return vec3u(s[o], s[o + 1], s[o + 2]); }
template <> WGB_DEV vec4u wgb_get<vec4u>(const u32* s, int o) { return vec4u(s[o], s[o + 1], s[o + 2], s[o + 3]); }
template <> WGB_DEV vec2i wgb_get<vec2i>(const u32* s, int o) { return vec2i((i32)s[o], (i32)s[o + 1]); }
template <> WGB_DEV vec3i wgb_get<vec3i>(const u32* s, int o) { return vec3i((i32)s[o], (i32)s[o + 1], (i32)s[o + 2]); }
template <> WGB_DEV vec4i wgb_get<vec4i>(const u32* s, int o) { return vec4i((i32)s[o], (i32)s[o + 1], (i32)s[o + 2], (i32)s[o + 3]); }
