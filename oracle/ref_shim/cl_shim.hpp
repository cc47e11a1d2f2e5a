// oracle/ref_shim/cl_shim.hpp -- TEST INFRASTRUCTURE: the OpenCL C the reference's solid-angle kernel uses
// (src/xmi_kernels.cl), spelled for a host C++ compiler: vector types with OpenCL's member names, the geometric built-ins,
// work-item queries fed by the driver's loop, address-space qualifiers as no-ops.  Floating-point semantics follow an
// OpenCL device with cl_khr_fp64 (the reference builds without -cl-single-precision-constant, src/xmi_solid_angle_cl.c:325):
// unsuffixed literals are double and promote the expressions they appear in.
#ifndef ORC_REF_SHIM_CL_HPP
#define ORC_REF_SHIM_CL_HPP
#include <math.h>
#include <limits.h>
#include <stddef.h>
#include <stdio.h>
#include <algorithm>

#define __kernel
#define __global
#define __constant const
#define __local
#define M_PI_F 3.14159274101257f

typedef unsigned int uint;
typedef unsigned long ulong;
typedef unsigned char uchar;

struct float3 {
	union { float x; float s0; };
	union { float y; float s1; };
	union { float z; float s2; };
};
static inline float3 make_float3(double a, double b, double c) { float3 f; f.x = (float)a; f.y = (float)b; f.z = (float)c; return f; }
static inline float3 operator+(float3 a, float3 b) { float3 f; f.x = a.x + b.x; f.y = a.y + b.y; f.z = a.z + b.z; return f; }
static inline float3 operator-(float3 a, float3 b) { float3 f; f.x = a.x - b.x; f.y = a.y - b.y; f.z = a.z - b.z; return f; }
static inline float3 operator*(float s, float3 a) { float3 f; f.x = s * a.x; f.y = s * a.y; f.z = s * a.z; return f; }
static inline float3 operator*(float3 a, float s) { return s * a; }
static inline float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline float length(float3 a) { return sqrtf(dot(a, a)); }

struct uint4 {
	union { uint x; uint s0; };
	union { uint y; uint s1; };
	union { uint z; uint s2; };
	union { uint w; uint s3; };
};

using std::max;
using std::min;

// work-item queries: set by the driver for every (tid0, tid1) it runs
static thread_local size_t orc_cl_gid[2], orc_cl_gsz[2];
static inline size_t get_global_id(int d) { return orc_cl_gid[d]; }
static inline size_t get_global_size(int d) { return orc_cl_gsz[d]; }
#endif
