/*
 * optb_demo.c -- the C ABI of include/optb.h driven from plain C, no Python and no torch:
 * a partially reflecting circular mirror at x = 5 (reflectivity 0.6, transmission 0.4) and a 5 x 5 monitor at
 * x = 8, four rays along +x. Every ray is popped three times (the ray itself, its reflection, its transmission)
 * and the transmitted part leaves one monitor row with intensity 0.4.
 *
 *   gcc -std=c99 -Iinclude examples/optb_demo.c -Loptable_b200 -loptb -Wl,-rpath,$PWD/optable_b200 -o optb_demo
 *
 * This is the table layout optable_b200/flatten.py produces from optable's object graph, written out by hand.
 */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "optb.h"

#define N 4

static void identity(double* m) {
  memset(m, 0, 9 * sizeof(double));
  m[0] = m[4] = m[8] = 1.0;
}

int main(void) {
  /* ---- scene tables ---- */
  int32_t node_i[OPTB_NI_STRIDE] = {0};
  double node_f[OPTB_NF_STRIDE] = {0};
  node_i[OPTB_NI_GEOM] = OPTB_G_CIRCLE;
  node_i[OPTB_NI_INTER] = OPTB_I_MIRROR;
  node_i[OPTB_NI_SKIP] = 1;
  node_i[OPTB_NI_CAPSLOT] = -1;
  node_i[OPTB_NI_ROCKIND] = OPTB_ROC_INF;
  node_i[OPTB_NI_LEAF] = 0;
  node_i[OPTB_NI_ORTHO] = 1;
  node_f[OPTB_NF_ORIGIN] = 5.0;
  identity(node_f + OPTB_NF_TINV);
  identity(node_f + OPTB_NF_T);
  node_f[OPTB_NF_P] = 1.0; /* radius */
  node_f[OPTB_NF_REFL] = 0.6;
  node_f[OPTB_NF_TRANS] = 0.4;
  int32_t mat_kind[1] = {OPTB_MAT_CONST};
  double mat_f[OPTB_MF_STRIDE] = {1.0};
  double mon_f[OPTB_MON_STRIDE] = {0};
  mon_f[OPTB_MON_ORIGIN] = 8.0;
  identity(mon_f + OPTB_MON_TINV);
  mon_f[OPTB_MON_HW] = mon_f[OPTB_MON_HH] = 2.5;
  mon_f[OPTB_MON_TY + 1] = 1.0;
  mon_f[OPTB_MON_TZ + 2] = 1.0;
  mon_f[OPTB_MON_ORTHO] = 1.0;
  double aux[1] = {0};
  optb_scene_desc desc;
  memset(&desc, 0, sizeof desc);
  desc.abi_version = OPTB_ABI_VERSION;
  desc.n_nodes = 1; desc.n_leaves = 1; desc.n_materials = 1; desc.n_monitors = 1; desc.n_capslots = 0; desc.n_aux = 1;
  desc.node_i = node_i; desc.node_f = node_f; desc.mat_kind = mat_kind; desc.mat_f = mat_f; desc.mon_f = mon_f; desc.aux = aux;

  /* ---- rays (SoA) ---- */
  double ox[N] = {0}, oy[N] = {-0.5, -0.2, 0.1, 0.4}, oz[N] = {0}, dx[N] = {1, 1, 1, 1}, dy[N] = {0}, dz[N] = {0};
  double inten[N] = {1, 1, 1, 1}, wl[N], qre[N] = {0}, qim[N] = {0}, pl[N] = {0}, nmed[N] = {1, 1, 1, 1};
  uint32_t flags[N] = {OPTB_RF_ALIVE, OPTB_RF_ALIVE, OPTB_RF_ALIVE, OPTB_RF_ALIVE};
  for (int i = 0; i < N; i++) wl[i] = 780e-7;
  optb_rays rays;
  memset(&rays, 0, sizeof rays);
  rays.n = N;
  rays.ox = ox; rays.oy = oy; rays.oz = oz; rays.dx = dx; rays.dy = dy; rays.dz = dz;
  rays.intensity = inten; rays.wavelength = wl; rays.q_re = qre; rays.q_im = qim; rays.pathlength = pl; rays.n_medium = nmed;
  rays.flags = flags;

  /* ---- results (host buffers) ---- */
  enum { SEGCAP = 64, HITCAP = 16 };
  static double seg[13][SEGCAP], hit[10][HITCAP];
  static uint32_t seg_flags[SEGCAP], seg_root[SEGCAP], seg_pop[SEGCAP], hit_root[HITCAP], hit_pop[HITCAP];
  static int32_t seg_leaf[SEGCAP], hit_monitor[HITCAP];
  static int64_t counters[OPTB_C_COUNT];
  optb_result res;
  memset(&res, 0, sizeof res);
  res.seg_capacity = SEGCAP; res.hit_capacity = HITCAP;
  double** segp = &res.seg_ox;
  for (int f = 0; f < 13; f++) segp[f] = seg[f];
  res.seg_flags = seg_flags; res.seg_root = seg_root; res.seg_pop = seg_pop; res.seg_leaf = seg_leaf;
  res.hit_monitor = hit_monitor; res.hit_root = hit_root; res.hit_pop = hit_pop;
  double** hitp = &res.hit_px;
  for (int f = 0; f < 10; f++) hitp[f] = hit[f];
  res.counters = counters;

  optb_params prm;
  memset(&prm, 0, sizeof prm);
  prm.max_trace_num = 2000; prm.unit = 1e-2; prm.record_segments = 1; prm.record_hits = 1; prm.n_families = N;

  optb_ctx* ctx = NULL;
  optb_scene* scene = NULL;
  int rc = optb_ctx_create(0, &ctx);
  if (rc) { fprintf(stderr, "optb_ctx_create: %d\n", rc); return 1; }
  rc = optb_scene_upload(ctx, &desc, &scene);
  if (!rc) rc = optb_trace_host(ctx, scene, &rays, &prm, &res);
  if (rc) { fprintf(stderr, "optb error %d: %s\n", rc, optb_last_error(ctx)); return 1; }

  printf("abi %d segments %lld interactions %lld monitor_rows %lld status %lld\n", optb_abi_version(),
         (long long)counters[OPTB_C_SEGMENTS], (long long)counters[OPTB_C_INTERACTIONS], (long long)counters[OPTB_C_HITS],
         (long long)counters[OPTB_C_STATUS]);
  int ok = counters[OPTB_C_SEGMENTS] == 3 * N && counters[OPTB_C_INTERACTIONS] == N && counters[OPTB_C_HITS] == N &&
           counters[OPTB_C_STATUS] == 0;
  for (int k = 0; k < (int)counters[OPTB_C_HITS]; k++) {
    const uint32_t r = hit_root[k];
    printf("row root %u pop %u monitor %d y %.17g intensity %.17g t %.17g\n", r, hit_pop[k], hit_monitor[k], hit[1][k], hit[3][k],
           hit[4][k]);
    ok = ok && r < N && fabs(hit[1][k] - oy[r]) < 1e-15 && fabs(hit[3][k] - 0.4) < 1e-15 && fabs(hit[4][k] - 3.0) < 1e-12;
  }
  for (int k = 0; k < (int)counters[OPTB_C_SEGMENTS]; k++)
    if (seg_pop[k] == 0) ok = ok && seg_leaf[k] == 0 && fabs(seg[6][k] - 5.0) < 1e-12; /* seg_length of the initial ray */
  optb_scene_destroy(ctx, scene);
  optb_ctx_destroy(ctx);
  printf(ok ? "OK\n" : "MISMATCH\n");
  return ok ? 0 : 2;
}
