/* oracle/ref_shim/gccfeatures.h -- TEST INFRASTRUCTURE: the compiler-feature macros Random123's compilerfeatures.h
 * expects from "gccfeatures.h" (the reference bundles only the OpenCL and Metal variants, src/Random123/).  Plain
 * portable choices: no SSE / AES-NI / inline-assembly paths -- Threefry4x32 needs none of them. */
#ifndef ORC_REF_SHIM_GCCFEATURES_H
#define ORC_REF_SHIM_GCCFEATURES_H
#include <assert.h>
#include <stdint.h>
#define R123_STATIC_INLINE static __inline__
#define R123_FORCE_INLINE(decl) decl __attribute__((always_inline))
#define R123_CUDA_DEVICE
#define R123_ASSERT(x) assert(x)
#define R123_BUILTIN_EXPECT(expr, likely) __builtin_expect(expr, likely)
#define R123_USE_CXX11_UNRESTRICTED_UNIONS 1
#define R123_USE_CXX11_STATIC_ASSERT 1
#define R123_USE_CXX11_CONSTEXPR 1
#define R123_USE_CXX11_EXPLICIT_CONVERSIONS 1
#define R123_USE_CXX11_RANDOM 0
#define R123_USE_CXX11_TYPE_TRAITS 1
#define R123_USE_CXX11_LONG_LONG 1
#define R123_USE_CXX11_STD_ARRAY 0
#define R123_USE_AES_NI 0
#define R123_USE_SSE4_2 0
#define R123_USE_SSE4_1 0
#define R123_USE_SSE 0
#define R123_USE_AES_OPENSSL 0
#define R123_USE_GNU_UINT128 1
#define R123_USE_ASM_GNU 0
#define R123_USE_CPUID_MSVC 0
#define R123_USE_X86INTRIN_H 0
#define R123_USE_IA32INTRIN_H 0
#define R123_USE_XMMINTRIN_H 0
#define R123_USE_EMMINTRIN_H 0
#define R123_USE_SMMINTRIN_H 0
#define R123_USE_WMMINTRIN_H 0
#define R123_USE_INTRIN_H 0
#define R123_USE_MULHILO32_ASM 0
#define R123_USE_MULHILO64_ASM 0
#define R123_USE_MULHILO64_MSVC_INTRIN 0
#define R123_USE_MULHILO64_CUDA_INTRIN 0
#define R123_USE_MULHILO64_OPENCL_INTRIN 0
#define R123_USE_MULHILO64_C99 0
#define R123_USE_MULHILO64_MULHI_INTRIN 0
#define R123_USE_MULHILO32_MULHI_INTRIN 0
#define R123_USE_PHILOX_64BIT 1
#define R123_USE_64BIT 1
#define R123_ULONG_LONG unsigned long long
#define R123_64BIT(x) x##ULL
#endif
