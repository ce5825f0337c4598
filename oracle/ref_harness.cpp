// ref_harness.cpp -- parameterised driver around the UNMODIFIED reference sources.
//
// TEST INFRASTRUCTURE ONLY (see oracle/rt_oracle.c header).  Nothing from the reference is
// copied: this translation unit #includes /root/reference/src/main.cpp where it lies (the
// include path is given by oracle/Makefile) with its main() renamed, so trace_ray, write_ppm,
// Scene, Camera and load_scene are the reference's own code, verbatim.  It exists because the
// reference hard-codes 1280x720 / depth 10 (src/main.cpp:95-97) and BASELINE.json's configs
// need other sizes.  Output binary: oracle/_ref/ref_harness (git-ignored).
//
// usage: ref_harness render <scene> <W> <H> <depth> <out.ppm> [omp]
//        ref_harness dump   <scene>                      (parsed scene, %.17g, one item per line)
//        ref_harness trace  <scene> <W> <H> <depth> <out.bin>   (hit-index / shadow-mask / counters)
#define main ref_main
#include "src/main.cpp"
#undef main

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace {

struct Counters { uint64_t closest = 0, hits = 0, shadow = 0, occluded = 0; };

// Counting mirror of src/main.cpp:16-58: same calls into the reference's Scene, plus
// bookkeeping.  Only used for the index/shadow dump, never for the image.
void mirror(const Ray &ray, const Scene &scene, int depth, int level, int32_t *hit, uint32_t *mask, Counters &c) {
  if (depth <= 0) return;
  double t; int idx;
  c.closest++;
  if (!scene.find_intersection(ray, t, idx)) { hit[level] = -1; return; }
  c.hits++;
  hit[level] = idx;
  Vec3 p = ray.origin + ray.direction * t;
  Vec3 n = scene.spheres[idx].normal_at(p);
  uint32_t m = 0;
  for (int l = 0; l < int(scene.lights.size()); l++) {
    c.shadow++;
    if (scene.in_shadow(p, scene.lights[l])) { c.occluded++; if (l < 32) m |= 1u << l; }
  }
  mask[level] = m;
  if (scene.spheres[idx].material.reflectivity > 0) {
    Vec3 rd = ray.direction - n * 2.0 * dot(ray.direction, n);
    Ray rr(p + n * EPSILON, rd);
    mirror(rr, scene, depth - 1, level + 1, hit, mask, c);
  }
}

}  // namespace

int main(int argc, char **argv) {
  if (argc < 3) { std::fprintf(stderr, "usage: see source header\n"); return 2; }
  std::string mode = argv[1];
  Scene scene = load_scene(argv[2]);
  if (mode == "dump") {
    std::printf("spheres %zu\n", scene.spheres.size());
    for (const Sphere &s : scene.spheres)
      std::printf("%.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", s.center.x, s.center.y, s.center.z, s.radius,
                  s.material.color.x, s.material.color.y, s.material.color.z, s.material.reflectivity, s.material.shininess);
    std::printf("lights %zu\n", scene.lights.size());
    for (const Light &l : scene.lights)
      std::printf("%.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", l.position.x, l.position.y, l.position.z, l.color.x,
                  l.color.y, l.color.z, l.intensity);
    std::printf("ambient %.17g %.17g %.17g\n", scene.ambient_light.x, scene.ambient_light.y, scene.ambient_light.z);
    std::printf("camera %.17g %.17g %.17g %.17g %.17g %.17g %.17g %d\n", scene.camera.position.x, scene.camera.position.y,
                scene.camera.position.z, scene.camera.look_at.x, scene.camera.look_at.y, scene.camera.look_at.z,
                scene.camera.fov, int(scene.has_camera));
    return 0;
  }
  if (argc < 7) { std::fprintf(stderr, "usage: see source header\n"); return 2; }
  const int W = std::atoi(argv[3]), H = std::atoi(argv[4]), D = std::atoi(argv[5]);
  Camera camera(scene.camera.position, scene.camera.look_at, scene.camera.fov);  // src/main.cpp:132
  if (mode == "render") {
    const bool use_omp = argc > 7 && std::string(argv[7]) == "omp";
    // optional bounded sample: only pixels whose linear index is a multiple of `step` (bench.py)
    const size_t step = argc > 8 ? (size_t)std::atoll(argv[8]) : 1;
    std::vector<Vec3> fb((size_t)W * H);
    auto t0 = std::chrono::high_resolution_clock::now();
    if (!use_omp) {
      for (int j = 0; j < H; j++)  // src/main.cpp:146-157
        for (int i = 0; i < W; i++) {
          if (step > 1 && ((size_t)j * W + i) % step) continue;
          double u = double(i) / (W - 1), v = double(j) / (H - 1);
          fb[(size_t)j * W + i] = trace_ray(camera.get_ray(u, v), scene, D);
        }
    } else {
#pragma omp parallel for schedule(dynamic) collapse(2)  // src/main.cpp:185
      for (int j = 0; j < H; j++)
        for (int i = 0; i < W; i++) {
          if (step > 1 && ((size_t)j * W + i) % step) continue;
          double u = double(i) / (W - 1), v = double(j) / (H - 1);
          fb[(size_t)j * W + i] = trace_ray(camera.get_ray(u, v), scene, D);
        }
    }
    std::chrono::duration<double> dt = std::chrono::high_resolution_clock::now() - t0;
    std::printf("%s time: %.6f seconds\n", use_omp ? "OpenMP" : "Serial", dt.count());
    if (std::strcmp(argv[6], "-") != 0) write_ppm(argv[6], fb, W, H);  // src/main.cpp:69-91
    return 0;
  }
  if (mode == "trace") {
    std::vector<int32_t> hit((size_t)W * H * D, -2);
    std::vector<uint32_t> mask((size_t)W * H * D, 0);
    Counters total;
#pragma omp parallel
    {
      Counters c;
#pragma omp for schedule(dynamic, 64) collapse(2)
      for (int j = 0; j < H; j++)
        for (int i = 0; i < W; i++) {
          double u = double(i) / (W - 1), v = double(j) / (H - 1);
          size_t p = (size_t)j * W + i;
          mirror(camera.get_ray(u, v), scene, D, 0, &hit[p * D], &mask[p * D], c);
        }
#pragma omp critical
      { total.closest += c.closest; total.hits += c.hits; total.shadow += c.shadow; total.occluded += c.occluded; }
    }
    FILE *f = std::fopen(argv[6], "wb");
    if (!f) return 1;
    int32_t hdr[4] = {W, H, D, 0};
    std::fwrite(hdr, sizeof(hdr), 1, f);
    uint64_t cn[4] = {total.closest, total.hits, total.shadow, total.occluded};
    std::fwrite(cn, sizeof(cn), 1, f);
    std::fwrite(hit.data(), sizeof(int32_t), hit.size(), f);
    std::fwrite(mask.data(), sizeof(uint32_t), mask.size(), f);
    std::fclose(f);
    std::printf("closest %llu hits %llu shadow %llu occluded %llu\n", (unsigned long long)total.closest,
                (unsigned long long)total.hits, (unsigned long long)total.shadow, (unsigned long long)total.occluded);
    return 0;
  }
  std::fprintf(stderr, "unknown mode %s\n", mode.c_str());
  return 2;
}
