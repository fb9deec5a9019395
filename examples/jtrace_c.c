/* jtrace_c.c -- a host in plain C: scenes/<name>/<name>.json -> image, through nothing but include/jtrace_b200.h.
 * The same phases as the reference's Jtrace.main (src/jtrace.jl:31-116): load scene, find camera, build bvh, make
 * lights, make state, `samples / batch` calls of trace_samples, save. Writes a binary PPM (P6) of the sRGB image.
 *
 *   gcc -O2 -Iinclude examples/jtrace_c.c -o jtrace_c -Ljulia-raytracer_b200 -ljtrace_b200 -Wl,-rpath,$PWD/julia-raytracer_b200
 *   ./jtrace_c scenes/cornellbox/cornellbox.json out.ppm [resolution] [samples] [sampler 1|2] [device-lights 0|1]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "jtrace_b200.h"

#define CHECK(call)                                                  \
  do {                                                               \
    int rc_ = (call);                                                \
    if (rc_ != JT_OK) {                                              \
      fprintf(stderr, "%s -> %d: %s\n", #call, rc_, jt_last_error()); \
      return 1;                                                      \
    }                                                                \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: %s scene.json out.ppm [resolution=640] [samples=16] [sampler=1] [device-lights=0]\n", argv[0]);
    return 2;
  }
  const int resolution = argc > 3 ? atoi(argv[3]) : 640, samples = argc > 4 ? atoi(argv[4]) : 16;
  const int sampler = argc > 5 ? atoi(argv[5]) : 1, device_lights = argc > 6 ? atoi(argv[6]) : 0;

  jt_host_scene* host = NULL;
  CHECK(jt_host_scene_load(argv[1], &host));              /* load_scene */
  for (int i = 0; i < jt_host_scene_num_notes(host); i++) printf("note: %s\n", jt_host_scene_note(host, i));
  int32_t camera = -1;
  CHECK(jt_host_scene_find_camera(host, "", &camera));    /* find_camera */
  CHECK(jt_host_scene_build(host, 0));                    /* make_scene_bvh + make_trace_lights */
  const jt_scene_desc* built = NULL;
  CHECK(jt_host_scene_desc(host, &built));
  jt_scene_desc desc = *built;
  jt_lights* lights = NULL;
  if (device_lights) {                                    /* the same CDFs, built on the GPU (N4) */
    CHECK(jt_lights_create(&desc, 0, 0, &lights));
    CHECK(jt_lights_desc(lights, &desc.lights, &desc.num_lights));
  }
  jt_scene* scene = NULL;
  CHECK(jt_scene_create(&desc, 0, &scene));               /* flatten + upload; the library keeps no host pointer */
  if (lights) jt_lights_destroy(lights);
  jt_host_scene_destroy(host);

  jt_params p;
  memset(&p, 0, sizeof(p));
  p.camera = camera; p.resolution = resolution; p.samples = samples; p.bounces = 8; p.sampler = sampler; p.clamp = 10;
  p.batch = 1; p.bvhstacksize = 128;
  jt_state* state = NULL;
  CHECK(jt_state_create(scene, &p, &state));              /* make_trace_state */
  int32_t width = 0, height = 0, done = 0;
  CHECK(jt_state_size(state, &width, &height, &done));
  while (done < p.samples) {                              /* the reference's loop, one call per batch */
    CHECK(jt_trace_samples(scene, state, &p));
    CHECK(jt_state_size(state, NULL, NULL, &done));
  }
  unsigned char* rgba = (unsigned char*)malloc((size_t)width * height * 4);
  if (!rgba) return 1;
  CHECK(jt_state_download_srgb8(state, rgba));            /* get_image + rgb_to_srgb + 8-bit quantisation on the GPU */
  jt_counters c;
  CHECK(jt_scene_counters(scene, &c, 0));
  FILE* f = fopen(argv[2], "wb");
  if (!f) return 1;
  fprintf(f, "P6\n%d %d\n255\n", width, height);
  for (size_t i = 0; i < (size_t)width * height; i++) fwrite(rgba + 4 * i, 1, 3, f);
  fclose(f);
  printf("%dx%d, %d spp: %llu camera paths, %llu scene rays, %llu light probes, %llu kernel launches -> %s\n", width, height,
         done, (unsigned long long)c.camera_paths, (unsigned long long)c.scene_rays, (unsigned long long)c.light_rays,
         (unsigned long long)c.kernel_launches, argv[2]);
  free(rgba);
  jt_state_destroy(state);
  jt_scene_destroy(scene);
  return 0;
}
