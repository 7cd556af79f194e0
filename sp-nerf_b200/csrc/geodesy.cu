// Ray generation and DSM point clouds: the two geometry steps either side of render_rays.
//
//  * spnerf_rays_from_geodetic replaces datasets/satellite_scene.py:38-68 (get_rays after the RPC localisation: geodetic
//    -> ECEF at the maximum / minimum altitude, origin, unit direction, near = 0, far = |far - near|), the float32 cast of
//    :65-66 and normalize_rays (:415-425), and can append the sun direction columns of :463-473, i.e. it writes the
//    (n, 11) rows render_rays consumes.  modules/utils.py:80-100 (geodetic_to_ecef, WGS-84) is evaluated in fp64 like
//    numpy does; the normalisation is float32 arithmetic on float32 values, like torch's in-place ops on a FloatTensor.
//  * spnerf_points_to_geodetic replaces datasets/satellite_scene.py:475-505 (get_latlonalt_from_nerf_prediction:
//    point = origin + direction * depth in fp64, de-normalised) + modules/utils.py:103-120 (ecef_to_latlon_custom).
//
// The RPC localisation itself (rpcm) and the UTM projection / rasterisation after it (pyproj, plyflatten) stay on the
// host: those libraries are not in this image, so nothing could be pinned against them.
// One thread per ray, coalesced loads; fp64 throughput is irrelevant at one transcendental chain per ray.
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/spnerf_b200.h"

namespace {

constexpr double kA = 6378137.0;            // WGS-84 semi-major axis (modules/utils.py:85)
constexpr double kB = 6356752.314245;       // semi-minor axis (:86)

// numpy evaluates every ufunc separately: products and sums round one by one (no fused multiply-add)
__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }

__device__ __forceinline__ void geodetic_to_ecef(double lat, double lon, double alt, double& x, double& y, double& z) {
  const double ratio = (kB * kB) / (kA * kA);                         // b ** 2 / a ** 2
  const double e2 = 1.0 - ratio;                                      // :87
  const double d2r = 3.14159265358979323846 / 180.0;                  // np.radians
  const double la = mul(lat, d2r), lo = mul(lon, d2r);
  const double sl = sin(la), cl = cos(la);
  const double N = kA / sqrt(sub(1.0, mul(e2, mul(sl, sl))));         // :93
  x = mul(mul(add(N, alt), cl), cos(lo));                             // :96-98
  y = mul(mul(add(N, alt), cl), sin(lo));
  z = mul(add(mul(ratio, N), alt), sl);
}

struct RaysP {
  const double* lon_near; const double* lat_near; const double* lon_far; const double* lat_far;
  double alt_near, alt_far;
  float cx, cy, cz, range;
  int normalize, has_sun;
  float sx, sy, sz;
  int64_t n; int stride;
  float* rays;
};

__global__ void rays_kernel(const RaysP p) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  double xn, yn, zn, xf, yf, zf;
  geodetic_to_ecef(p.lat_near[i], p.lon_near[i], p.alt_near, xn, yn, zn);          // satellite_scene.py:42-44
  geodetic_to_ecef(p.lat_far[i], p.lon_far[i], p.alt_far, xf, yf, zf);              // :47-49
  const double dx = xf - xn, dy = yf - yn, dz = zf - zn;                            // :55
  const double norm = sqrt(add(add(mul(dx, dx), mul(dy, dy)), mul(dz, dz)));        // np.linalg.norm
  float o0 = (float)xn, o1 = (float)yn, o2 = (float)zn;                             // :65-66 float32 cast
  const float d0 = (float)(dx / norm), d1 = (float)(dy / norm), d2 = (float)(dz / norm);
  float near = 0.f, far = (float)norm;                                              // :60-61
  if (p.normalize) {                                                                // :415-425, float32 in-place ops
    o0 = __fdiv_rn(__fsub_rn(o0, p.cx), p.range);
    o1 = __fdiv_rn(__fsub_rn(o1, p.cy), p.range);
    o2 = __fdiv_rn(__fsub_rn(o2, p.cz), p.range);
    near = __fdiv_rn(near, p.range);
    far = __fdiv_rn(far, p.range);
  }
  float* r = p.rays + i * p.stride;
  r[0] = o0; r[1] = o1; r[2] = o2; r[3] = d0; r[4] = d1; r[5] = d2; r[6] = near; r[7] = far;
  if (p.has_sun) { r[8] = p.sx; r[9] = p.sy; r[10] = p.sz; }
}

struct PointsP {
  const float* rays; int stride; const float* depth;
  double cx, cy, cz, range;
  int64_t n;
  double* lat; double* lon; double* alt;
};

__global__ void points_kernel(const PointsP p) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  const float* r = p.rays + i * p.stride;
  const double dep = (double)p.depth[i];
  // satellite_scene.py:488-499: rays.double(), o + d * depth, * range, + centre
  const double x = add(mul(add((double)r[0], mul((double)r[3], dep)), p.range), p.cx);
  const double y = add(mul(add((double)r[1], mul((double)r[4], dep)), p.range), p.cy);
  const double z = add(mul(add((double)r[2], mul((double)r[5], dep)), p.range), p.cz);
  // modules/utils.py:107-120
  const double a = 6378137.0, e = 8.1819190842622e-2;
  const double asq = a * a, esq = e * e;
  const double b = sqrt(mul(asq, 1.0 - esq)), bsq = mul(b, b);
  const double ep = sqrt(sub(asq, bsq) / bsq);
  const double pp = sqrt(add(mul(x, x), mul(y, y)));
  const double th = atan2(mul(a, z), mul(b, pp));
  const double sth = sin(th), cth = cos(th);
  const double lon = atan2(y, x);
  // np.sin(th) ** 3 is pow(sin, 3) in numpy (only ** 2 is special-cased to a product)
  const double lat = atan2(add(z, mul(mul(mul(ep, ep), b), pow(sth, 3.0))), sub(pp, mul(mul(esq, a), pow(cth, 3.0))));
  const double sla = sin(lat);
  const double N = a / sqrt(sub(1.0, mul(esq, mul(sla, sla))));
  p.alt[i] = sub(pp / cos(lat), N);
  p.lon[i] = mul(lon, 180.0) / 3.14159265358979323846;
  p.lat[i] = mul(lat, 180.0) / 3.14159265358979323846;
}

}  // namespace

extern "C" int spnerf_rays_from_geodetic(const SpnerfRaysFromGeodetic* a, void* stream) {
  if (!a || !a->lon_near || !a->lat_near || !a->lon_far || !a->lat_far || !a->rays) return SPNERF_ERR_BAD_ARG;
  if (a->n_rays <= 0) return a->n_rays == 0 ? 0 : SPNERF_ERR_BAD_ARG;
  if (a->row_stride < (a->has_sun ? 11 : 8)) return SPNERF_ERR_BAD_ARG;
  if (a->normalize && !(a->range > 0.f)) return SPNERF_ERR_BAD_ARG;
  RaysP p;
  p.lon_near = a->lon_near; p.lat_near = a->lat_near; p.lon_far = a->lon_far; p.lat_far = a->lat_far;
  p.alt_near = a->alt_near; p.alt_far = a->alt_far;
  p.cx = a->center[0]; p.cy = a->center[1]; p.cz = a->center[2]; p.range = a->range;
  p.normalize = a->normalize; p.has_sun = a->has_sun;
  p.sx = a->sun_dir[0]; p.sy = a->sun_dir[1]; p.sz = a->sun_dir[2];
  p.n = a->n_rays; p.stride = a->row_stride; p.rays = a->rays;
  rays_kernel<<<(unsigned)((p.n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}

extern "C" int spnerf_points_to_geodetic(const SpnerfPointsToGeodetic* a, void* stream) {
  if (!a || !a->rays || !a->depth || !a->lat || !a->lon || !a->alt) return SPNERF_ERR_BAD_ARG;
  if (a->n_rays <= 0) return a->n_rays == 0 ? 0 : SPNERF_ERR_BAD_ARG;
  if (a->row_stride < 6) return SPNERF_ERR_BAD_ARG;
  PointsP p;
  p.rays = a->rays; p.stride = a->row_stride; p.depth = a->depth;
  p.cx = (double)a->center[0]; p.cy = (double)a->center[1]; p.cz = (double)a->center[2]; p.range = (double)a->range;
  p.n = a->n_rays; p.lat = a->lat; p.lon = a->lon; p.alt = a->alt;
  points_kernel<<<(unsigned)((p.n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}
