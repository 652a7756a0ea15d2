// optb_flags.cuh -- the ambiguity mask of SURVEY A.9 ("the stated epsilon" of the parity bar), evaluated on the device.
//
// A diagnostics pass (params.flag_ambiguity): after the closest-hit search of a pop has picked its winner, the pop is
// walked once more the reference's way (every box, every leaf whose boxes pass) and every decision of the reference
// that hinges on less than the stated margin sets a bit: near-tie of the two closest surfaces, a hit point within
// 1e-9 * scale of an aperture edge, grazing incidence, sin_t within 1e-9 of total internal reflection, t within
// 1e-11 of the 1e-9 self-intersection guard or at the ray's length limit, a sign-scan sample with |f| < 1e-12 or a
// bracket at the t = 100 clip, a box test decided by less than 1e-11 or with | |d_axis| - 1e-8 | < 1e-10.
// The pass is separate from the hot search on purpose: it is allowed to be slow, the search is not.
#pragma once
#include "optb_device.cuh"

namespace optb {

OPTB_DEV bool near(double x, double eps) { return fabs(x) <= eps; }

// general slab test with margins (solver.py:5-48); returns hit, sets SLAB in `amb` when the decision is marginal:
// an axis within 1e-10 of the np.isclose "parallel" threshold, a parallel axis whose origin sits on a box face, or the
// entry parameter of one axis within 1e-11 of the (1e-12-slackened) exit parameter of ANOTHER axis. Entry and exit of
// the same axis are not compared: for a zero-thickness box (the lab box of an unrotated planar leaf) they are the same
// number computed twice, which is what the reference's 1e-12 slack is there for, and nothing hinges on rounding.
OPTB_DEV bool flag_slab(double ox, double oy, double oz, double dx, double dy, double dz, const double* __restrict__ bb,
                        double& t1o, double& t2o, unsigned& amb) {
  double t1 = 0.0, t2 = INFINITY;
  const double o[3] = {ox, oy, oz}, d[3] = {dx, dy, dz};
  double tn[3], tf[3];
  bool par[3];
#pragma unroll
  for (int ax = 0; ax < 3; ax++) {
    const double bmin = bb[2 * ax], bmax = bb[2 * ax + 1];
    if (near(fabs(d[ax]) - 1e-8, 1e-10)) amb |= OPTB_AMB_SLAB;
    par[ax] = fabs(d[ax]) <= 1e-8;
    tn[ax] = 0.0; tf[ax] = INFINITY;
    if (par[ax]) {
      const double scale = 1e-11 * fmax(1.0, fmax(fabs(bmin), fabs(bmax)));
      if (near(o[ax] - bmin, scale) || near(o[ax] - bmax, scale)) amb |= OPTB_AMB_SLAB;
      if (o[ax] < bmin || o[ax] > bmax) { t1 = 1.0; t2 = 0.0; }
    } else {
      const double inv = 1.0 / d[ax];
      const double ta = (bmin - o[ax]) * inv, tb = (bmax - o[ax]) * inv;
      tn[ax] = fmin(ta, tb); tf[ax] = fmax(ta, tb);
      t1 = fmax(t1, tn[ax]);
      t2 = fmin(t2, tf[ax]);
    }
  }
#pragma unroll
  for (int ax = 0; ax < 3; ax++)
#pragma unroll
    for (int bx = 0; bx < 3; bx++)
      if (ax != bx && !par[ax] && !par[bx] && near(tf[bx] + 1e-12 - tn[ax], 1e-11)) amb |= OPTB_AMB_SLAB;
  t1o = t1; t2o = t2;
  return (t2 + 1e-12 >= t1) && (t2 >= 0.0);
}

// distance of a 2-D point to the nearest edge of a polygon record (surfaces.py:534-558)
OPTB_DEV double poly_edge_distance(const double* __restrict__ rec, double Px, double Py, double Pz) {
  const int nv = (int)rec[0];
  const double* v = rec + OPTB_POLY_HEADER;
  const double ex = Px - rec[4], ey = Py - rec[5], ez = Pz - rec[6];
  const double px = dot3(ex, ey, ez, rec[7], rec[8], rec[9]);
  const double py = dot3(ex, ey, ez, rec[10], rec[11], rec[12]);
  double best = INFINITY;
  for (int i = 0; i < nv; i++) {
    const int j = (i + 1 == nv) ? 0 : i + 1;
    const double x1 = v[2 * i], y1 = v[2 * i + 1], x2 = v[2 * j], y2 = v[2 * j + 1];
    const double ux = x2 - x1, uy = y2 - y1, L2 = ux * ux + uy * uy;
    double s = L2 > 0 ? ((px - x1) * ux + (py - y1) * uy) / L2 : 0.0;
    s = fmin(fmax(s, 0.0), 1.0);
    const double qx = x1 + s * ux - px, qy = y1 + s * uy - py;
    best = fmin(best, sqrt(qx * qx + qy * qy));
  }
  return best;
}

// is the local point P within `eps * scale` of the edge of a simple planar shape?
OPTB_DEV bool planar_edge(const SceneView& sv, int kind, double p0, double p1, double Px, double Py, double Pz, double eps) {
  if (kind == OPTB_G_CIRCLE) return near(sqrt(dot3(Px, Py, Pz, Px, Py, Pz)) - p0, eps * fmax(p0, 1.0));
  if (kind == OPTB_G_RECT) {
    const double ey = fabs(Py) - p0, ez = fabs(Pz) - p1, m = eps * fmax(1.0, fmax(p0, p1));
    return (near(ey, m) && ez <= m) || (near(ez, m) && ey <= m);
  }
  if (kind == OPTB_G_POLY2D) return poly_edge_distance(sv.aux + (long long)p0, Px, Py, Pz) <= 2e-9;
  return false;
}

// One pop, the reference's way, with margins. best_node / best_t: what the search decided (-1: nothing hit).
OPTB_DEV unsigned flag_pop(const SceneView& sv, const Ray& ray, int best_node, double best_t, double unit) {
  unsigned amb = 0u;
  if (!(ray.flags & OPTB_RF_ALIVE)) return 0u;
  const double eps = 1e-9;
  const double t_rel = (best_node >= 0) ? best_t * (1.0 + eps) : INFINITY;  // candidates beyond cannot matter
  int i = 0;
  const int n = sv.n_nodes;
  // A box test decided by less than the margin only matters if something under that box could be hit: such boxes are
  // entered either way, and the SLAB bit is set when a leaf below them yields a candidate (a ray leaving a surface
  // starts ON that surface's box -- t2 = +-1 ulp -- and the surface's own leaf test then fails t >= 1e-9: no bit).
  int marginal_until = 0;
  while (i < n) {
    const double* tv = sv.trav + i * 8;
    const int2 gs = *reinterpret_cast<const int2*>(tv + 6);
    if (*reinterpret_cast<const int*>(tv + 7)) {
      double t1, t2;
      unsigned box_amb = 0u;
      const bool hit = flag_slab(ray.ox, ray.oy, ray.oz, ray.dx, ray.dy, ray.dz, tv, t1, t2, box_amb);
      const bool marginal = box_amb != 0u || near(t2, 1e-11);  // (t2 ~ 0: the ray starts on the box)
      if (marginal) marginal_until = max(marginal_until, gs.y);
      else if (!hit) { i = gs.y; continue; }
    }
    const int g = gs.x;
    const int cur = i++;
    if (g == OPTB_G_GROUP || g == OPTB_G_GRID) continue;
    const int32_t* ni = sv.ni + cur * OPTB_NI_STRIDE;
    const double* nf = sv.nf + cur * OPTB_NF_STRIDE;
    const double* p = nf + OPTB_NF_P;
    double ox, oy, oz, dx, dy, dz;
    to_local(nf + OPTB_NF_ORIGIN, nf + OPTB_NF_TINV, ray, ni[OPTB_NI_ORTHO] != 0, ox, oy, oz, dx, dy, dz);
    double t = -1.0;
    bool edge = false;
    if (is_planar_kind(g)) {
      if (dx == 0.0) continue;
      const double tp = -ox / dx;
      if (near(tp - 1e-9, 1e-11)) amb |= OPTB_AMB_EPS;
      if (isfinite(ray.len) && near(tp - ray.len, eps * fmax(ray.len, 1e-3))) amb |= OPTB_AMB_EPS;
      if (!(tp >= 1e-9) || tp > ray.len) continue;
      const double Px = fma(tp, dx, ox), Py = fma(tp, dy, oy), Pz = fma(tp, dz, oz);
      bool in;
      if (g == OPTB_G_CSG) {
        in = csg_within(sv, ni, p, Px, Py, Pz);
        if ((int)p[0] != 2) {
          edge = planar_edge(sv, (int)p[1], p[2], p[3], Px, Py, Pz, eps) || planar_edge(sv, (int)p[4], p[5], p[6], Px, Py, Pz, eps);
        } else {  // nested composite: near the edge of any of its shapes
          const double* prog = sv.aux + ni[OPTB_NI_AUX];
          for (int k = 0; k < (int)prog[0]; k++)
            if ((int)prog[1 + 3 * k] > 0) edge = edge || planar_edge(sv, (int)prog[1 + 3 * k], prog[2 + 3 * k], prog[3 + 3 * k], Px, Py, Pz, eps);
        }
      } else if (g == OPTB_G_POLY2D) {
        in = poly_within(sv.aux + ni[OPTB_NI_AUX], Px, Py, Pz);
        edge = poly_edge_distance(sv.aux + ni[OPTB_NI_AUX], Px, Py, Pz) <= 2e-9;
      } else {
        in = planar_within(sv, g, p[0], p[1], Px, Py, Pz);
        edge = planar_edge(sv, g, p[0], p[1], Px, Py, Pz, eps);
      }
      if (edge && tp <= t_rel) amb |= OPTB_AMB_APERTURE;
      if (!in) continue;
      t = tp;
    } else {
      // curved branch: the bracket of intersect_point_local :197-233 with its margins, then the real root search
      double bb[6];
      if (g == OPTB_G_SPHERE) { for (int k = 0; k < 6; k++) bb[k] = p[2 + k]; }
      else if (g == OPTB_G_ASPHERE) { bb[0] = p[6]; bb[1] = p[7]; bb[2] = -p[0]; bb[3] = p[0]; bb[4] = -p[0]; bb[5] = p[0]; }
      else if (g == OPTB_G_CYL) { bb[0] = -p[0]; bb[1] = p[0]; bb[2] = -p[0]; bb[3] = p[0]; bb[4] = -p[1] / 2; bb[5] = p[1] / 2; }
      else { const double* rec = sv.aux + ni[OPTB_NI_AUX]; for (int k = 0; k < 6; k++) bb[k] = rec[13 + k]; }
      double t1, t2;
      unsigned loc_amb = 0u;
      flag_slab(ox, oy, oz, dx, dy, dz, bb, t1, t2, loc_amb);
      if (t2 + 1e-9 < t1) { if (near(t2 + 1e-9 - t1, 1e-11)) amb |= OPTB_AMB_SLAB; continue; }
      if (near(t2 - 100.0, 1e-7)) amb |= OPTB_AMB_SCAN;
      t1 = fmax(t1, 0.0); t2 = fmin(t2, 100.0);
      const double a = t1 - 1e-9, b = t2 + 1e-9, step = (b - a) / 9.0;
      for (int k = 0; k < 10; k++) {  // |f| at a sample point (true f of the surface, surfaces.py)
        const double ts = sample_t(k, a, b, step);
        const double Px = fma(ts, dx, ox), Py = fma(ts, dy, oy), Pz = fma(ts, dz, oz);
        double f;
        if (g == OPTB_G_SPHERE) f = sqrt(dot3(Px, Py, Pz, Px, Py, Pz)) - p[0];
        else if (g == OPTB_G_CYL) f = sqrt(fma(Px, Px, Py * Py)) - p[0];
        else if (g == OPTB_G_ASPHERE) f = Px + f_asphere(ni[OPTB_NI_AUX], p + 1, sqrt(fma(Py, Py, Pz * Pz)));
        else { const double* rec = sv.aux + ni[OPTB_NI_AUX]; f = dot3(rec[1], rec[2], rec[3], Px - rec[4], Py - rec[5], Pz - rec[6]); }
        // the scan right after leaving this very surface starts ON it (sample 0 at t = -1e-9): not a coincidence
        if (fabs(f) < 1e-12 && !(k == 0 && a < 0.0)) amb |= OPTB_AMB_SCAN;
      }
      t = intersect_leaf<true>(sv, ni, nf, ox, oy, oz, dx, dy, dz, ray.len, INFINITY);
      if (t >= 0.0) {
        if (near(t - 1e-9, 1e-11)) amb |= OPTB_AMB_EPS;
        if (isfinite(ray.len) && near(t - ray.len, eps * fmax(ray.len, 1e-3))) amb |= OPTB_AMB_EPS;
        const double Px = fma(t, dx, ox), Py = fma(t, dy, oy), Pz = fma(t, dz, oz);
        if (g == OPTB_G_SPHERE) edge = near(Px - (p[0] - p[1]), eps * fmax(fabs(p[0]), 1.0)) || near(Px - p[0], 1e-12);
        else if (g == OPTB_G_ASPHERE) edge = near(sqrt(fma(Py, Py, Pz * Pz)) - p[0], eps * fmax(p[0], 1.0));
        else if (g == OPTB_G_CYL) {
          const double th = atan2(Py, Px);
          edge = near(th - p[2], eps) || near(th - p[3], eps) || near(fabs(Pz) - p[1] / 2, eps * fmax(p[1], 1.0));
        } else edge = poly_edge_distance(sv.aux + ni[OPTB_NI_AUX], Px, Py, Pz) <= 2e-9;
        if (edge && t <= t_rel) amb |= OPTB_AMB_APERTURE;
      }
      if (!(t >= 0.0)) continue;
    }
    if (cur < marginal_until && t <= t_rel) amb |= OPTB_AMB_SLAB;  // a candidate that exists only if a marginal box passes
    // a second surface as close as the winner (the reference decides by the last bit of t; ties go to list order)
    if (best_node >= 0 && cur != best_node && near(t - best_t, eps * fmax(fabs(best_t), 1e-3))) amb |= OPTB_AMB_TIE;
  }
  if (best_node >= 0) {  // physics margins at the winner: grazing incidence, the TIR threshold
    const int32_t* ni = sv.ni + best_node * OPTB_NI_STRIDE;
    const double* nf = sv.nf + best_node * OPTB_NF_STRIDE;
    const int kind = ni[OPTB_NI_INTER];
    if (kind == OPTB_I_MIRROR || kind == OPTB_I_REFRACT) {
      double ox, oy, oz, dx, dy, dz, nx, ny, nz, roc;
      to_local(nf + OPTB_NF_ORIGIN, nf + OPTB_NF_TINV, ray, ni[OPTB_NI_ORTHO] != 0, ox, oy, oz, dx, dy, dz);
      surf_normal<true>(sv, ni, nf, fma(best_t, dx, ox), fma(best_t, dy, oy), fma(best_t, dz, oz), nx, ny, nz, roc);
      const double dn = dot3(dx, dy, dz, nx, ny, nz);
      if (fabs(dn) < 1e-6) amb |= OPTB_AMB_GRAZING;
      if (kind == OPTB_I_REFRACT) {
        const double wl_m = ray.wl * unit;
        const double n1 = material_n(sv, ni[OPTB_NI_MAT1], wl_m), n2 = material_n(sv, ni[OPTB_NI_MAT2], wl_m);
        const double nin = dn < 0 ? n1 : n2, nout = dn < 0 ? n2 : n1;
        const double ci = fmin(fmax(dn, -1.0), 1.0);
        const double sin_t = nin * sqrt(1.0 - ci * ci) / nout;
        if (near(sin_t - 1.0, 1e-9)) amb |= OPTB_AMB_TIR;
      }
    }
  }
  return amb;
}

}  // namespace optb
