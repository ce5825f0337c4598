// scene_io.cpp -- host-side scene file reader and PPM writer behind the C ABI.
//
// Re-statement (not a copy) of the reference's text formats, which are the API surface the
// north star keeps: include/scene_loader.h:27-135 (scene grammar, warn-and-skip behaviour,
// "Loaded scene:" line) and src/main.cpp:69-91 (P3 writer).  No GPU is needed for these.
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "rt_internal.h"

struct rt_scene {
  std::vector<double> spheres;  // N x 10, file column order
  std::vector<double> lights;   // L x 7
  double ambient[3];
  double camera[7];
  int has_camera;
};

namespace {

// Mimics `istream >> double` (libstdc++ num_get): skip whitespace, accept
// [+-]digits[.digits][(e|E)[+-]digits]; a dangling exponent marker makes the extraction fail
// (the stream would have consumed it), anything else stops the token without consuming it.
bool extract_double(const char *&p, double &out) {
  while (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\v' || *p == '\f' || *p == '\r') ++p;
  const char *s = p;
  const char *q = p;
  if (*q == '+' || *q == '-') ++q;
  const char *digits0 = q;
  while (*q >= '0' && *q <= '9') ++q;
  size_t nd = (size_t)(q - digits0);
  if (*q == '.') {
    ++q;
    const char *f0 = q;
    while (*q >= '0' && *q <= '9') ++q;
    nd += (size_t)(q - f0);
  }
  if (nd == 0) return false;
  if (*q == 'e' || *q == 'E') {
    const char *e = q + 1;
    if (*e == '+' || *e == '-') ++e;
    if (!(*e >= '0' && *e <= '9')) return false;
    while (*e >= '0' && *e <= '9') ++e;
    q = e;
  }
  std::string tok(s, (size_t)(q - s));
  char *end = nullptr;
  out = std::strtod(tok.c_str(), &end);
  if (end == tok.c_str()) return false;
  // libstdc++'s num_get (`iss >> double`, include/scene_loader.h:63-101) sets failbit when strtod overflows to
  // +-HUGE_VAL, so the reference warns and skips such a line; underflow is accepted there and here
  if (out == HUGE_VAL || out == -HUGE_VAL) return false;
  p = q;
  return true;
}

bool extract_n(const char *&p, double *dst, int n) {
  for (int i = 0; i < n; i++)
    if (!extract_double(p, dst[i])) return false;
  return true;
}

}  // namespace

extern "C" int rt_scene_load(const char *path, int verbose, rt_scene **out) {
  if (!path || !out) return rt_fail(RT_ERR_ARG, "rt_scene_load: NULL argument");
  FILE *f = std::fopen(path, "rb");
  if (!f) return rt_fail(RT_ERR_IO, std::string("Could not open scene file: ") + path);
  rt_scene *sc = new (std::nothrow) rt_scene();
  if (!sc) { std::fclose(f); return rt_fail(RT_ERR_NOMEM, "rt_scene_load: out of memory"); }
  // include/scene.h:22,31 defaults: Vec3() ambient, camera (0,0,0)->(0,0,-1) fov 60
  sc->ambient[0] = sc->ambient[1] = sc->ambient[2] = 0.0;
  const double cam0[7] = {0, 0, 0, 0, 0, -1, 60.0};
  std::memcpy(sc->camera, cam0, sizeof(cam0));
  sc->has_camera = 0;

  std::string line;
  int line_number = 0;
  bool eof = false;
  while (!eof) {
    line.clear();
    int ch;
    bool got_any = false;
    while ((ch = std::fgetc(f)) != EOF) {
      got_any = true;
      if (ch == '\n') break;
      line.push_back((char)ch);
    }
    if (ch == EOF) { eof = true; if (!got_any) break; }
    line_number++;
    if (line.empty() || line[0] == '#') continue;              // scene_loader.h:43-45
    size_t start = line.find_first_not_of(" \t");               // :48-52
    if (start == std::string::npos) continue;
    const char *p = line.c_str() + start;
    if (*p == '#') continue;                                    // :55-57
    // `iss >> type`: leading whitespace already gone; token runs to the next whitespace
    const char *t0 = p;
    while (*p && !(*p == ' ' || *p == '\t' || *p == '\n' || *p == '\v' || *p == '\f' || *p == '\r')) ++p;
    std::string type(t0, (size_t)(p - t0));
    double v[10];
    if (type == "sphere") {
      if (!extract_n(p, v, 10)) { std::fprintf(stderr, "Warning: Invalid sphere at line %d, skipping\n", line_number); continue; }
      sc->spheres.insert(sc->spheres.end(), v, v + 10);
    } else if (type == "light") {
      if (!extract_n(p, v, 7)) { std::fprintf(stderr, "Warning: Invalid light at line %d, skipping\n", line_number); continue; }
      sc->lights.insert(sc->lights.end(), v, v + 7);
    } else if (type == "ambient") {
      if (!extract_n(p, v, 3)) { std::fprintf(stderr, "Warning: Invalid ambient at line %d, skipping\n", line_number); continue; }
      std::memcpy(sc->ambient, v, 3 * sizeof(double));
    } else if (type == "camera") {
      if (!extract_n(p, v, 7)) { std::fprintf(stderr, "Warning: Invalid camera at line %d, skipping\n", line_number); continue; }
      std::memcpy(sc->camera, v, 7 * sizeof(double));
      sc->has_camera = 1;
    } else {
      std::fprintf(stderr, "Warning: Unknown type '%s' at line %d, skipping\n", type.c_str(), line_number);
    }
  }
  std::fclose(f);
  if (verbose) {
    std::printf("Loaded scene: %zu spheres, %zu lights\n", sc->spheres.size() / 10, sc->lights.size() / 7);
    std::fflush(stdout);
  }
  *out = sc;
  return RT_OK;
}

extern "C" int rt_scene_counts(const rt_scene *s, int *nspheres, int *nlights, int *has_camera) {
  if (!s) return rt_fail(RT_ERR_ARG, "rt_scene_counts: NULL scene");
  if (nspheres) *nspheres = (int)(s->spheres.size() / 10);
  if (nlights) *nlights = (int)(s->lights.size() / 7);
  if (has_camera) *has_camera = s->has_camera;
  return RT_OK;
}

extern "C" int rt_scene_data(const rt_scene *s, const double **spheres, const double **lights,
                             const double **ambient, const double **camera) {
  if (!s) return rt_fail(RT_ERR_ARG, "rt_scene_data: NULL scene");
  if (spheres) *spheres = s->spheres.data();
  if (lights) *lights = s->lights.data();
  if (ambient) *ambient = s->ambient;
  if (camera) *camera = s->camera;
  return RT_OK;
}

extern "C" void rt_scene_free(rt_scene *s) { delete s; }

// P3 writer: one pass over a 256-entry table of pre-formatted decimal strings into a big
// buffer (the reference streams through ofstream<<double, ~24 MB of text at 1080p).
extern "C" int rt_write_ppm(const char *path, const uint8_t *rgb, int width, int height) {
  if (!path || !rgb || width < 1 || height < 1) return rt_fail(RT_ERR_ARG, "rt_write_ppm: bad argument");
  FILE *f = std::fopen(path, "wb");
  if (!f) return rt_fail(RT_ERR_IO, std::string("Could not open output file: ") + path);
  char tab[256][4];
  unsigned char len[256];
  for (int i = 0; i < 256; i++) len[i] = (unsigned char)std::snprintf(tab[i], 4, "%d", i);
  std::fprintf(f, "P3\n%d %d\n255\n", width, height);
  std::vector<char> buf((size_t)width * 12 + 16);
  for (int j = height - 1; j >= 0; j--) {
    const uint8_t *row = rgb + (size_t)j * width * 3;
    char *o = buf.data();
    for (int i = 0; i < width; i++) {
      for (int c = 0; c < 3; c++) {
        unsigned v = row[i * 3 + c];
        std::memcpy(o, tab[v], len[v]);
        o += len[v];
        *o++ = (c == 2) ? '\n' : ' ';
      }
    }
    if (std::fwrite(buf.data(), 1, (size_t)(o - buf.data()), f) != (size_t)(o - buf.data())) {
      std::fclose(f);
      return rt_fail(RT_ERR_IO, "rt_write_ppm: short write");
    }
  }
  if (std::fclose(f) != 0) return rt_fail(RT_ERR_IO, "rt_write_ppm: close failed");
  return RT_OK;
}
