/*
 * rt_oracle.c -- CPU restatement (FP64, plain C) of the reference's serial ray tracer.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (cs420-ray-tracer_b200/csrc) never links, imports or calls anything in oracle/.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement byte-for-byte
 * against PPMs produced by the unmodified reference sources (oracle/_ref, built by
 * oracle/Makefile from /root/reference in place) and against the md5s recorded in
 * BASELINE.md for simple/medium/complex at 1280x720 d10 and 1920x1080 d5.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference).  All arithmetic is double, evaluated in the reference's operation
 * order; build with -ffp-contract=off so no FMA is ever formed (the reference is built
 * for baseline x86-64, which has none).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* include/ray_math_constants.h:22-23 ; include/scene.h:38 */
static const double RTO_EPSILON = 0.001;
static const double RTO_INFINITY = 1e20;
static const double RTO_K_SPECULAR = 0.5;

typedef struct { double x, y, z; } v3;

/* include/vec3.h:13-17 */
static inline v3 v3_make(double x, double y, double z) { v3 r = {x, y, z}; return r; }
static inline v3 v3_add(v3 a, v3 b) { return v3_make(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v3_sub(v3 a, v3 b) { return v3_make(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v3_mulv(v3 a, v3 b) { return v3_make(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 v3_scale(v3 a, double t) { return v3_make(a.x * t, a.y * t, a.z * t); }
/* include/vec3.h:19-20 */
static inline double v3_length(v3 a) { return sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
static inline v3 v3_normalized(v3 a) { double len = v3_length(a); return v3_make(a.x / len, a.y / len, a.z / len); }
/* include/vec3.h:23-25 */
static inline double v3_dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
/* include/vec3.h:27-29 */
static inline v3 v3_cross(v3 a, v3 b) {
    return v3_make(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
/* include/vec3.h:31-33 : v - n * 2.0 * dot(v, n)  ==  v - ((n*2.0) * dot) */
static inline v3 v3_reflect(v3 v, v3 n) { return v3_sub(v, v3_scale(v3_scale(n, 2.0), v3_dot(v, n))); }

/* std::max / std::min exactly as libstdc++ defines them (matters only for NaN) */
static inline double std_max(double a, double b) { return (a < b) ? b : a; }
static inline double std_min(double a, double b) { return (b < a) ? b : a; }

/* include/ray.h:6-15 : the constructor normalises the direction */
typedef struct { v3 origin, direction; } ray_t;
static inline ray_t ray_make(v3 o, v3 d) { ray_t r; r.origin = o; r.direction = v3_normalized(d); return r; }

/* include/sphere.h:8-20 ; file column order: cx cy cz r R G B metallic roughness shininess
 * (include/scene_loader.h:62-84: metallic -> reflectivity, roughness dropped) */
typedef struct { v3 center; double radius; v3 color; double reflectivity, shininess; } sphere_t;
/* include/scene.h:10-14 ; file order: x y z R G B intensity (intensity never used) */
typedef struct { v3 position, color; double intensity; } light_t;

typedef struct {
    const sphere_t *spheres; int nspheres;
    const light_t *lights; int nlights;
    v3 ambient;
} scene_t;

typedef struct {
    uint64_t closest_queries;   /* find_intersection calls issued by trace_ray  (R_c) */
    uint64_t hits;              /* of those, how many hit                              */
    uint64_t shadow_queries;    /* in_shadow calls                               (R_s) */
    uint64_t occluded;          /* of those, how many returned true                    */
    uint64_t alive[32];         /* closest queries per bounce level                    */
} rto_counters;

/* include/sphere.h:26-59 */
static inline int sphere_intersect(const sphere_t *s, const ray_t *ray, double *t) {
    v3 oc = v3_sub(ray->origin, s->center);
    double a = v3_dot(ray->direction, ray->direction);
    double b = 2.0 * v3_dot(oc, ray->direction);
    double c = v3_dot(oc, oc) - s->radius * s->radius;
    double discriminant = b * b - 4 * a * c;
    if (discriminant < 0) return 0;
    if (discriminant == 0) { *t = -b / (2 * a); return 1; }
    double t1 = (-b - sqrt(discriminant)) / (2 * a);
    double t2 = (-b + sqrt(discriminant)) / (2 * a);
    if (std_max(t1, t2) < 0) return 0;
    *t = std_min(t1, t2);
    if (*t < 0) *t = std_max(t1, t2);
    return 1;
}

/* include/scene.h:41-61 : strict '<', ascending index => lowest index wins ties */
static inline int find_intersection(const scene_t *sc, const ray_t *ray, double *t, int *idx) {
    *t = RTO_INFINITY;
    *idx = -1;
    for (int i = 0; i < sc->nspheres; i++) {
        double tt = 0;
        if (sphere_intersect(&sc->spheres[i], ray, &tt)) {
            if (tt < *t) { *idx = i; *t = tt; }
        }
    }
    return *idx >= 0;
}

/* include/scene.h:65-86 */
static inline int in_shadow(const scene_t *sc, v3 point, const light_t *light) {
    v3 to_light = v3_sub(light->position, point);
    double light_distance = v3_length(to_light);
    v3 light_dir = v3_normalized(to_light);
    ray_t shadow_ray = ray_make(v3_add(point, v3_scale(light_dir, RTO_EPSILON)), light_dir);
    double t; int idx;
    if (find_intersection(sc, &shadow_ray, &t, &idx)) return t < light_distance;
    return 0;
}

/* include/scene.h:89-121 ; *shadow_mask gets bit l set when light l is occluded */
static inline v3 shade(const scene_t *sc, v3 point, v3 normal, const sphere_t *mat, v3 view_dir,
                       uint32_t *shadow_mask, rto_counters *cnt) {
    v3 color = v3_mulv(sc->ambient, mat->color);
    uint32_t mask = 0;
    for (int l = 0; l < sc->nlights; l++) {
        if (cnt) cnt->shadow_queries++;
        if (in_shadow(sc, point, &sc->lights[l])) {
            if (l < 32) mask |= (1u << l);
            if (cnt) cnt->occluded++;
            continue;
        }
        v3 light_dir = v3_normalized(v3_sub(sc->lights[l].position, point));
        double n_dot_l = std_max(0.0, v3_dot(normal, light_dir));
        v3 diffuse = v3_scale(v3_scale(mat->color, 1.0 - mat->reflectivity), n_dot_l);
        v3 reflect_dir = v3_reflect(v3_scale(light_dir, -1), normal);
        double r_dot_v = std_max(0.0, v3_dot(reflect_dir, view_dir));
        double spec_factor = pow(r_dot_v, mat->shininess);
        v3 specular = v3_scale(v3_scale(sc->lights[l].color, RTO_K_SPECULAR), spec_factor);
        color = v3_add(v3_add(specular, diffuse), color);
    }
    if (shadow_mask) *shadow_mask = mask;
    return color;
}

/* src/main.cpp:16-58 ; level = max_depth - depth (0 for the camera ray).
 * hit_idx / shadow_mask (optional) are per-pixel arrays of max_depth entries. */
static v3 trace_ray(const ray_t *ray, const scene_t *sc, int depth, int level,
                    int32_t *hit_idx, uint32_t *shadow_mask, rto_counters *cnt) {
    if (depth <= 0) return v3_make(0, 0, 0);
    double t; int idx;
    if (cnt) { cnt->closest_queries++; if (level < 32) cnt->alive[level]++; }
    if (!find_intersection(sc, ray, &t, &idx)) {
        if (hit_idx) hit_idx[level] = -1;
        double ts = 0.5 * (ray->direction.y + 1.0);
        return v3_add(v3_scale(v3_make(1, 1, 1), 1.0 - ts), v3_scale(v3_make(0.5, 0.7, 1.0), ts));
    }
    if (cnt) cnt->hits++;
    if (hit_idx) hit_idx[level] = idx;
    const sphere_t *s = &sc->spheres[idx];
    v3 hit = v3_add(ray->origin, v3_scale(ray->direction, t));
    v3 norm = v3_normalized(v3_sub(hit, s->center));               /* include/sphere.h:62-64 */
    v3 view_dir = v3_normalized(v3_sub(ray->origin, hit));
    v3 col = shade(sc, hit, norm, s, view_dir, shadow_mask ? &shadow_mask[level] : NULL, cnt);
    if (s->reflectivity > 0) {
        v3 reflected_dir = v3_sub(ray->direction, v3_scale(v3_scale(norm, 2.0), v3_dot(ray->direction, norm)));
        ray_t reflected = ray_make(v3_add(hit, v3_scale(norm, RTO_EPSILON)), reflected_dir);
        v3 rc = trace_ray(&reflected, sc, depth - 1, level + 1, hit_idx, shadow_mask, cnt);
        double refl = s->reflectivity;
        col = v3_add(v3_scale(col, 1.0 - refl), v3_scale(rc, refl));
    }
    return col;
}

/* include/camera.h:10-15 */
typedef struct { v3 position, forward, right, up; double fov; } camera_t;
static camera_t camera_make(v3 pos, v3 look_at, double fov) {
    camera_t c; c.position = pos; c.fov = fov;
    c.forward = v3_normalized(v3_sub(look_at, pos));
    c.right = v3_normalized(v3_cross(c.forward, v3_make(0, 1, 0)));
    c.up = v3_normalized(v3_cross(c.right, c.forward));
    return c;
}
/* include/camera.h:17-25 */
static inline ray_t camera_get_ray(const camera_t *c, double u, double v) {
    double aspect = 1.0;
    double scale = tan(c->fov * 0.5 * M_PI / 180.0);
    v3 direction = v3_add(v3_add(c->forward, v3_scale(c->right, (u - 0.5) * scale * aspect)),
                          v3_scale(c->up, (v - 0.5) * scale));
    return ray_make(c->position, v3_normalized(direction));
}

/* src/main.cpp:84-86 */
static inline uint8_t quantise(double c) { return (uint8_t)(int)(255.99 * std_min(1.0, c)); }

/*
 * Render.  spheres: N x 10 doubles in scene-file column order, lights: L x 7.
 * Outputs (each may be NULL):
 *   fb          W*H*3 doubles, row j=0 = bottom of the image (src/main.cpp:153-154)
 *   rgb         W*H*3 bytes, same row order, quantised as src/main.cpp:84-86
 *   hit_idx     W*H*max_depth int32: sphere index per bounce level, -1 = miss, -2 = level not traced
 *   shadow_mask W*H*max_depth uint32: bit l = light l occluded at that level's hit
 *   cnt         ray counters (summed over all pixels rendered)
 * pix_step > 1 renders only pixels whose linear index p = j*W+i satisfies p % pix_step == 0
 * (deterministic subsample for very large configs); others are left untouched.
 * nthreads <= 0 -> all OpenMP threads; 1 -> serial (src/main.cpp:146-157), else the
 * schedule(dynamic) collapse(2) loop of src/main.cpp:185-199.
 */
/* Sample position inside the pixel, in pixels (default 0,0 = the reference's serial renderer).  The 2x2
 * supersampling of the reference's ray_cuda -a (src/main_gpu.cu:253-256) renders the offsets (0,0) (.5,0)
 * (0,.5) (.5,.5): u = (i + off_x)/(W-1), v = (j + off_y)/(H-1).  Set before rto_render, reset after. */
static double g_off_x = 0.0, g_off_y = 0.0;
void rto_set_sample_offset(double off_x, double off_y) { g_off_x = off_x; g_off_y = off_y; }

int rto_render(const double *spheres, int N, const double *lights, int L, const double *ambient,
               const double *cam_pos, const double *cam_look, double fov_deg,
               int W, int H, int max_depth,
               double *fb, uint8_t *rgb, int32_t *hit_idx, uint32_t *shadow_mask,
               rto_counters *cnt, int pix_step, int nthreads) {
    if (W < 1 || H < 1 || N < 0 || L < 0) return -1;
    sphere_t *sp = (sphere_t *)malloc(sizeof(sphere_t) * (size_t)(N > 0 ? N : 1));
    light_t *li = (light_t *)malloc(sizeof(light_t) * (size_t)(L > 0 ? L : 1));
    if (!sp || !li) { free(sp); free(li); return -2; }
    for (int i = 0; i < N; i++) {
        const double *r = spheres + (size_t)i * 10;
        sp[i].center = v3_make(r[0], r[1], r[2]); sp[i].radius = r[3];
        sp[i].color = v3_make(r[4], r[5], r[6]); sp[i].reflectivity = r[7]; sp[i].shininess = r[9];
    }
    for (int i = 0; i < L; i++) {
        const double *r = lights + (size_t)i * 7;
        li[i].position = v3_make(r[0], r[1], r[2]); li[i].color = v3_make(r[3], r[4], r[5]); li[i].intensity = r[6];
    }
    scene_t sc; sc.spheres = sp; sc.nspheres = N; sc.lights = li; sc.nlights = L;
    sc.ambient = v3_make(ambient[0], ambient[1], ambient[2]);
    camera_t cam = camera_make(v3_make(cam_pos[0], cam_pos[1], cam_pos[2]),
                               v3_make(cam_look[0], cam_look[1], cam_look[2]), fov_deg);
    if (pix_step < 1) pix_step = 1;
    if (cnt) memset(cnt, 0, sizeof(*cnt));
    int md = max_depth > 0 ? max_depth : 0;
#ifdef _OPENMP
    int nt = nthreads > 0 ? nthreads : omp_get_max_threads();
#else
    int nt = 1;
#endif
    rto_counters total; memset(&total, 0, sizeof(total));
#pragma omp parallel num_threads(nt)
    {
        rto_counters local; memset(&local, 0, sizeof(local));
#pragma omp for schedule(dynamic, 64) collapse(2)
        for (int j = 0; j < H; j++) {
            for (int i = 0; i < W; i++) {
                size_t p = (size_t)j * W + i;
                if (p % (size_t)pix_step) continue;
                double u = ((double)i + g_off_x) / (W - 1);
                double v = ((double)j + g_off_y) / (H - 1);
                ray_t ray = camera_get_ray(&cam, u, v);
                int32_t *hi = hit_idx ? hit_idx + p * (size_t)md : NULL;
                uint32_t *sm = shadow_mask ? shadow_mask + p * (size_t)md : NULL;
                if (hi) for (int k = 0; k < md; k++) hi[k] = -2;
                if (sm) for (int k = 0; k < md; k++) sm[k] = 0;
                v3 c = trace_ray(&ray, &sc, max_depth, 0, hi, sm, cnt ? &local : NULL);
                if (fb) { fb[p * 3 + 0] = c.x; fb[p * 3 + 1] = c.y; fb[p * 3 + 2] = c.z; }
                if (rgb) { rgb[p * 3 + 0] = quantise(c.x); rgb[p * 3 + 1] = quantise(c.y); rgb[p * 3 + 2] = quantise(c.z); }
            }
        }
#pragma omp critical
        {
            total.closest_queries += local.closest_queries; total.hits += local.hits;
            total.shadow_queries += local.shadow_queries; total.occluded += local.occluded;
            for (int k = 0; k < 32; k++) total.alive[k] += local.alive[k];
        }
    }
    if (cnt) *cnt = total;
    free(sp); free(li);
    return 0;
}

/* Single-ray known-answer entry (populi-files/demo1_intersection.cpp:36-40 vectors):
 * returns hit flag, writes t.  dir is normalised as the Ray ctor does. */
int rto_intersect(const double *origin, const double *dir, const double *center, double radius, double *t) {
    sphere_t s; memset(&s, 0, sizeof(s));
    s.center = v3_make(center[0], center[1], center[2]); s.radius = radius;
    ray_t r = ray_make(v3_make(origin[0], origin[1], origin[2]), v3_make(dir[0], dir[1], dir[2]));
    double tt = 0; int h = sphere_intersect(&s, &r, &tt); *t = tt; return h;
}

/* src/main.cpp:69-91 : P3 text, rows top (j=H-1) to bottom.  rgb is bottom-row-first. */
int rto_write_ppm(const char *path, const uint8_t *rgb, int W, int H) {
    FILE *f = fopen(path, "w");
    if (!f) return -1;
    fprintf(f, "P3\n%d %d\n255\n", W, H);
    for (int j = H - 1; j >= 0; j--)
        for (int i = 0; i < W; i++) {
            const uint8_t *p = rgb + ((size_t)j * W + i) * 3;
            fprintf(f, "%d %d %d\n", p[0], p[1], p[2]);
        }
    return fclose(f);
}

int rto_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
