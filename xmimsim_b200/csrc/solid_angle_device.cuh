// solid_angle_device.cuh -- one point of the solid-angle Monte Carlo (xmi_single_solid_angle_calculation,
// src/xmi_solid_angle_f.F90:432-710), shared by the grid kernel (solid_angle.cu) and by the history kernel's
// fallback for interaction points outside the grid (xmi_get_solid_angle, :783-789).
#pragma once
#include "cuda_util.cuh"

// (double)w * 2^-32 without an integer-to-double conversion (quarter-rate pipe): the word is placed in the mantissa of
// 2^52, the subtraction is exact -- the same value as xmb_u01(w).
__device__ __forceinline__ double u01_exact(uint32_t w) {
	return (__hiloint2double(0x43300000, (int)w) - 4503599627370496.0) * (1.0 / 4294967296.0);
}

struct SaDetector {   // detector window + collimator, detector frame (src/xmi_main.F90:1784-1836)
	int collimator_present;
	double detector_radius, collimator_radius, collimator_height;
};

struct SaCone {       // sampling cone of one (r, theta) point
	double st, ct;          // sin / cos of the cone axis elevation (rotation matrix, :593-597)
	double one_m_cos;       // 1 - cos(apex)
	double cone_sa;         // 2 pi (1 - cos apex)
	double py, pz, hz;      // photon_line%point = (0, py, pz); hz = collimator_height - pz
	bool outside;           // the point is above the collimator opening: rays must pass it first
	bool dead;              // shadowed by a conical collimator: solid angle 0 (:533-536)
};

// cone selection: no / cylindrical / conical collimator (:481-558), apex (:569-590)
__device__ __forceinline__ SaCone sa_cone_setup(const SaDetector &D, double r1, double theta1) {
	SaCone c;
	double s1, c1;
	sincos(theta1, &s1, &c1);
	double r, theta, base_radius;
	c.outside = false; c.dead = false;
	if (!D.collimator_present) {
		r = r1; theta = theta1; base_radius = D.detector_radius;
	} else if (fabs(D.collimator_radius - D.detector_radius) < 0.000001) {
		if (r1 * c1 <= D.detector_radius) { r = r1; theta = theta1; base_radius = D.detector_radius; }
		else {
			r = sqrt(r1 * r1 - 2.0 * r1 * s1 * D.collimator_height + D.collimator_height * D.collimator_height);
			theta = acos(r1 * c1 / r);
			base_radius = D.collimator_radius;
		}
		c.outside = r1 * s1 > D.collimator_height;
	} else {
		if (r1 * c1 <= D.detector_radius &&
		    r1 * s1 <= D.collimator_height * (r1 * c1 - D.detector_radius) / (D.collimator_radius - D.detector_radius)) {
			r = r1; theta = theta1; base_radius = D.detector_radius;
		} else if (r1 * s1 <= D.collimator_height) {
			c.dead = true; r = r1; theta = theta1; base_radius = D.detector_radius;
		} else {
			r = sqrt(r1 * r1 - 2.0 * r1 * s1 * D.collimator_height + D.collimator_height * D.collimator_height);
			theta = acos(r1 * c1 / r);
			base_radius = D.collimator_radius;
		}
		c.outside = r1 * s1 > D.collimator_height;
	}
	sincos(theta, &c.st, &c.ct);
	const double beta = atan(base_radius / r);
	double alpha1 = atan(base_radius * c.st / (r - base_radius * c.ct));
	if (alpha1 <= 0.0) alpha1 += M_PI;
	const double cos_apex = cos(fmax(beta, alpha1));
	c.cone_sa = 2 * M_PI * (1.0 - cos_apex);
	c.one_m_cos = 1.0 - cos_apex;
	c.py = r1 * c1; c.pz = r1 * s1;
	c.hz = D.collimator_height - c.pz;
	return c;
}

// One ray.  The reference draws theta = acos(1 - u1 (1 - cos apex)), phi = 2 pi u2 and takes sin/cos of both
// (:633-650).  The same direction without the inverse: cos theta = 1 - t, sin theta = sqrt(t (2 - t)) with
// t = u1 (1 - cos apex) (exact identity, better conditioned for narrow cones), (sin, cos) phi by sincospi(2 u2).
// The two plane intersections (:667-691) are tested multiplied through by dz^2 > 0 -- no division:
// x dz = (h - pz) dx, y dz = (h - pz) dy + py dz for the plane z = h.
__device__ __forceinline__ bool sa_ray_hits(const SaCone &c, double det_r2, double col_r2, uint32_t w1, uint32_t w2) {
	const double t = u01_exact(w1) * c.one_m_cos;
	const double cth = 1.0 - t, sth = sqrt(t * (2.0 - t));
	double sph, cph;
	sincospi(2.0 * u01_exact(w2), &sph, &cph);
	const double dx = sth * cph, cy = sth * sph;
	// MATMUL(rotation_matrix, dirv_from_cone), rows (1,0,0), (0,-sin,-cos), (0,cos,-sin)
	const double dy = -c.st * cy - c.ct * cth, dz = c.ct * cy - c.st * cth;
	const double dz2 = dz * dz;
	const double a = c.pz * dx, b = c.py * dz - c.pz * dy;
	bool hit = dz < 0.0 && a * a + b * b <= det_r2 * dz2;
	if (c.outside) {
		const double a2 = c.hz * dx, b2 = c.hz * dy + c.py * dz;
		hit = hit && a2 * a2 + b2 * b2 <= col_r2 * dz2;
	}
	return hit;
}

// rays 2p and 2p+1 of a point: one Philox block (as the OpenCL kernel's Threefry use, src/xmi_kernels.cl:395-406)
__device__ __forceinline__ int sa_pair_hits(const SaCone &c, double det_r2, double col_r2, uint4 rnd, bool second) {
	int h = sa_ray_hits(c, det_r2, col_r2, rnd.x, rnd.y) ? 1 : 0;
	h += (second && sa_ray_hits(c, det_r2, col_r2, rnd.z, rnd.w)) ? 1 : 0;
	return h;
}
