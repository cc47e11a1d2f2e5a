/* A stand-in libxrl for tests/test_xraylib_binding_cpu.py: exports the xraylib 4 symbols xmb_xrl_from_library binds,
 * with the xraylib 4 signatures (trailing xrl_error **), returning values that encode their arguments so that the
 * test can check every thunk forwards the right ones.  Not physics. */
#include <stddef.h>
#define E_ void **error
double AtomicWeight(int Z, E_) { (void)error; return 2.0 * Z; }
double EdgeEnergy(int Z, int s, E_) { (void)error; return Z + 0.01 * s; }
double LineEnergy(int Z, int l, E_) { (void)error; return Z - 0.001 * l; }
double FluorYield(int Z, int s, E_) { (void)error; return 0.001 * Z + 0.01 * s; }
double RadRate(int Z, int l, E_) { (void)error; return 0.5 - 0.001 * l + 1e-6 * Z; }
double CosKronTransProb(int Z, int t, E_) { (void)error; return 0.01 * t + 1e-5 * Z; }
double JumpFactor(int Z, int s, E_) { (void)error; return 1.0 + s + 0.01 * Z; }
double CS_Total_Kissel(int Z, double E, E_) { (void)error; return 1000.0 * Z + E; }
double CS_Photo_Total(int Z, double E, E_) { (void)error; return 900.0 * Z + E; }
double CS_Photo_Partial(int Z, int s, double E, E_) { (void)error; return 100.0 * Z + 10.0 * s + E; }
double CS_Rayl(int Z, double E, E_) { (void)error; return 10.0 * Z + E; }
double CS_Compt(int Z, double E, E_) { (void)error; return 20.0 * Z + E; }
double FF_Rayl(int Z, double q, E_) { (void)error; return Z - q; }
double SF_Compt(int Z, double q, E_) { (void)error; return Z * q; }
double ComptonProfile(int Z, double pz, E_) { (void)error; return Z / (1.0 + pz); }
double ElectronConfig_Biggs(int Z, int s, E_) { (void)error; return Z + s; }
double ComptonProfile_Partial(int Z, int s, double pz, E_) { (void)error; return Z + s + pz; }
/* P<shell>_<mode>_kissel: value = tag + sum of the handed-over cross sections weighted by position, so a wrong
 * argument count or order changes it */
#define P0(name, tag) double name(int Z, double E, E_) { (void)error; return tag + Z + E; }
#define P1(name, tag) double name(int Z, double E, double a, E_) { (void)error; return tag + Z + E + 2 * a; }
#define P2(name, tag) double name(int Z, double E, double a, double b, E_) { (void)error; return tag + Z + E + 2 * a + 3 * b; }
#define P3(name, tag) double name(int Z, double E, double a, double b, double c, E_) { (void)error; return tag + Z + E + 2 * a + 3 * b + 4 * c; }
#define P4(name, tag) double name(int Z, double E, double a, double b, double c, double d, E_) { (void)error; return tag + Z + E + 2 * a + 3 * b + 4 * c + 5 * d; }
#define P5(name, tag) double name(int Z, double E, double a, double b, double c, double d, double e, E_) { (void)error; return tag + Z + E + 2 * a + 3 * b + 4 * c + 5 * d + 6 * e; }
#define P6(name, tag) double name(int Z, double E, double a, double b, double c, double d, double e, double f, E_) { (void)error; return tag + Z + E + 2 * a + 3 * b + 4 * c + 5 * d + 6 * e + 7 * f; }
#define P7(name, tag) double name(int Z, double E, double a, double b, double c, double d, double e, double f, double g, E_) { (void)error; return tag + Z + E + 2 * a + 3 * b + 4 * c + 5 * d + 6 * e + 7 * f + 8 * g; }
#define P8(name, tag) double name(int Z, double E, double a, double b, double c, double d, double e, double f, double g, double h, E_) { (void)error; return tag + Z + E + 2 * a + 3 * b + 4 * c + 5 * d + 6 * e + 7 * f + 8 * g + 9 * h; }
P0(PL1_pure_kissel, 1e-3) P1(PL2_pure_kissel, 2e-3) P2(PL3_pure_kissel, 3e-3)
P0(PM1_pure_kissel, 4e-3) P1(PM2_pure_kissel, 5e-3) P2(PM3_pure_kissel, 6e-3) P3(PM4_pure_kissel, 7e-3) P4(PM5_pure_kissel, 8e-3)
#define CASC(mode, t) \
	P1(PL1_##mode##_kissel, t + 1e-3) P2(PL2_##mode##_kissel, t + 2e-3) P3(PL3_##mode##_kissel, t + 3e-3) \
	P4(PM1_##mode##_kissel, t + 4e-3) P5(PM2_##mode##_kissel, t + 5e-3) P6(PM3_##mode##_kissel, t + 6e-3) \
	P7(PM4_##mode##_kissel, t + 7e-3) P8(PM5_##mode##_kissel, t + 8e-3)
CASC(auger_cascade, 0.1) CASC(rad_cascade, 0.2) CASC(full_cascade, 0.3)
