// exact_fp64.cuh -- the reference's FP64 geometry, operation for operation, for the device.
//
// Every arithmetic step uses the round-to-nearest intrinsics (__dadd_rn, __dmul_rn, ...),
// which nvcc never contracts into FMAs, so results are bit-identical to the reference built
// with g++ -O3 for baseline x86-64 (no FMA instructions): IEEE-754 double add/mul/div/sqrt are
// correctly rounded on both sides.  These routines DECIDE hits, sphere indices and shadow
// booleans; the FP32 code elsewhere only narrows down which spheres they have to look at.
//
// Reference lines restated (paths relative to the reference root):
//   include/vec3.h:13-33      Vec3 algebra, dot, cross, reflect
//   include/ray.h:12          Ray ctor normalises the direction
//   include/sphere.h:26-64    Sphere::intersect / normal_at
//   include/scene.h:65-86     Scene::in_shadow (shadow-ray construction)
//   include/camera.h:17-25    Camera::get_ray
//   src/main.cpp:32-48        hit point, normal, reflected ray
#ifndef RT_EXACT_FP64_CUH
#define RT_EXACT_FP64_CUH

#include <cuda_runtime.h>

namespace rtx {

struct d3 { double x, y, z; };

#define RT_DF __device__ __forceinline__

RT_DF double dadd(double a, double b) { return __dadd_rn(a, b); }
RT_DF double dsub(double a, double b) { return __dsub_rn(a, b); }
RT_DF double dmul(double a, double b) { return __dmul_rn(a, b); }
RT_DF double ddiv(double a, double b) { return __ddiv_rn(a, b); }
RT_DF double dsqrt(double a) { return __dsqrt_rn(a); }

RT_DF d3 mk(double x, double y, double z) { d3 r; r.x = x; r.y = y; r.z = z; return r; }
RT_DF d3 add(d3 a, d3 b) { return mk(dadd(a.x, b.x), dadd(a.y, b.y), dadd(a.z, b.z)); }
RT_DF d3 sub(d3 a, d3 b) { return mk(dsub(a.x, b.x), dsub(a.y, b.y), dsub(a.z, b.z)); }
RT_DF d3 scale(d3 a, double t) { return mk(dmul(a.x, t), dmul(a.y, t), dmul(a.z, t)); }
// include/vec3.h:23-25 : (ax*bx + ay*by) + az*bz
RT_DF double dot(d3 a, d3 b) { return dadd(dadd(dmul(a.x, b.x), dmul(a.y, b.y)), dmul(a.z, b.z)); }
// include/vec3.h:19-20
RT_DF double length(d3 a) { return dsqrt(dot(a, a)); }
RT_DF d3 normalized(d3 a) { double len = length(a); return mk(ddiv(a.x, len), ddiv(a.y, len), ddiv(a.z, len)); }

// std::max / std::min as libstdc++ defines them
RT_DF double std_max(double a, double b) { return (a < b) ? b : a; }
RT_DF double std_min(double a, double b) { return (b < a) ? b : a; }

// include/sphere.h:26-59.  `a` = dot(d,d) of the ray (hoisted: it does not depend on the
// sphere), r2 = radius*radius (one rounding, precomputed on the host in double).
RT_DF bool intersect(d3 o, d3 d, double a, d3 c, double r2, double &t) {
  d3 oc = sub(o, c);
  double b = dmul(2.0, dot(oc, d));
  double cc = dsub(dot(oc, oc), r2);
  double disc = dsub(dmul(b, b), dmul(dmul(4.0, a), cc));
  if (disc < 0) return false;
  double two_a = dmul(2.0, a);
  if (disc == 0) { t = ddiv(-b, two_a); return true; }
  double s = dsqrt(disc);
  double t1 = ddiv(dsub(-b, s), two_a);
  double t2 = ddiv(dadd(-b, s), two_a);
  if (std_max(t1, t2) < 0) return false;
  t = std_min(t1, t2);
  if (t < 0) t = std_max(t1, t2);
  return true;
}

// src/main.cpp:32 and include/sphere.h:62-64
RT_DF d3 hit_point(d3 o, d3 d, double t) { return add(o, scale(d, t)); }
RT_DF d3 normal_at(d3 p, d3 c) { return normalized(sub(p, c)); }

// src/main.cpp:45-48 : reflected ray; returns origin/direction as the Ray ctor stores them
RT_DF void reflect_ray(d3 d, d3 hit, d3 n, double eps, d3 &o2, d3 &d2) {
  d3 rd = sub(d, scale(scale(n, 2.0), dot(d, n)));
  o2 = add(hit, scale(n, eps));
  d2 = normalized(rd);
}

// include/scene.h:70-76 : shadow ray towards a light; ldist is measured from the un-offset point
RT_DF void shadow_ray(d3 p, d3 lpos, double eps, d3 &o2, d3 &d2, double &ldist) {
  d3 to_light = sub(lpos, p);
  ldist = length(to_light);
  d3 ldir = normalized(to_light);
  o2 = add(p, scale(ldir, eps));
  d2 = normalized(ldir);
}

// include/camera.h:17-25 with (u-0.5)*scale*aspect and (v-0.5)*scale precomputed per column /
// row on the host (same double operations, same order); direction normalised twice.
RT_DF d3 camera_dir(d3 fwd, d3 right, d3 up, double su, double sv) {
  d3 dir = add(add(fwd, scale(right, su)), scale(up, sv));
  return normalized(normalized(dir));
}

}  // namespace rtx
#endif
