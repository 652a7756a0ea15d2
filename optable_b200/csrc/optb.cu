// optb.cu -- C-ABI CUDA library (include/optb.h) for the optable bounce loop on B200 (sm_100a).
//
// Kernels:
//   trace_kernel      whole bounce loop. One thread owns one ray of the current wavefront; warps pull 32-ray
//                     chunks from a global work counter (persistent grid). The scene tables are staged into
//                     shared memory with a TMA bulk copy (cp.async.bulk + mbarrier). A root whose alive set
//                     is a single ray keeps bouncing in registers (closest hit -> physics -> monitors ->
//                     next bounce); split interactions park their children in a sparse slot pair.
//   tile_sums/scan_sums/slots     ordered, index-only stream compaction: the list of occupied child slots in
//                     reference BFS order (which the pop cap of optical_table.py:86-97 needs); rays are not moved.
//   mark_kernel       first/last wavefront index of every root (rank of a ray inside its root's generation).
//   sort_prep + cub radix sort    processing order of a generation by coherence key (storage order stays BFS).
// Segments and monitor rows are appended with warp-aggregated atomics and carry their (root, pop) key.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <cub/device/device_radix_sort.cuh>
#include <stddef.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <new>
#include <vector>

#include "optb_device.cuh"
#include "optb_flags.cuh"

using namespace optb;

namespace {

#ifndef OPTB_MIN_BLOCKS
#define OPTB_MIN_BLOCKS 4
#endif
#ifndef OPTB_GRID_WALK
#define OPTB_GRID_WALK 1   // 0: ignore the lattice descriptors (A/B switch; OPTB_G_GRID nodes are then plain groups)
#endif
#ifndef OPTB_BLOCK
#define OPTB_BLOCK 128
#endif
#ifndef OPTB_LANE_REFILL
#define OPTB_LANE_REFILL 1   // 0: a warp works through one 32-ray chunk at a time in every variant (A/B switch)
#endif
#ifndef OPTB_BIG_EXTRA
#define OPTB_BIG_EXTRA 1   // extra resident CTAs per SM for the variants whose scene tables stay in L1/L2
#endif
#ifndef OPTB_CULL_FAR
#define OPTB_CULL_FAR 1   // 0: boxes are only tested the reference's way, never used to dismiss by distance (A/B switch)
#endif
#ifndef OPTB_VOTE_SMEM
#define OPTB_VOTE_SMEM 0   // 1: the walk of shared-memory scenes also re-converges its lanes with a vote before each leaf test
#endif
#ifndef OPTB_PARK_ALWAYS
#define OPTB_PARK_ALWAYS 0   // 1: every variant parks the radiometric ray state in shared memory during the hit search
                             // (0: only asphere / lattice variants; measured r2u: c4 115.0 -> 110.8 ms, c3 8.00 -> 7.81)
#endif
#ifndef OPTB_WIN_BATCH
#define OPTB_WIN_BATCH 1   // lattice windows of up to 8 x 8 cells: the cells' box tests as one batch, two at a time
                           // (ripa 21.45 -> 21.19 ms per 1e6 rays; 0 = one candidate at a time, A/B switch)
#endif
#ifndef OPTB_PHASE_SYNC
#define OPTB_PHASE_SYNC 1   // 0: no per-pop CTA barrier in any variant (A/B switch)
#endif
constexpr int kBlock = OPTB_BLOCK;
constexpr int kResident(int ctas_of_128) { return ctas_of_128 * 128 / kBlock > 0 ? ctas_of_128 * 128 / kBlock : 1; }
constexpr int kTile = 2048;  // children-scan tile (entries per block)
constexpr int kScanBlock = 256;
constexpr int kRayF64 = 13;

// Ray storage in the workspace (wavefront W and sparse children C): one 128-byte record per ray. The wavefront is read
// through two indirections (coherence order -> dense BFS index -> sparse slot), i.e. by gathers: with a record per ray a
// gather touches exactly four full 32-byte sectors, where one column per field touched seventeen sectors for the same
// 128 bytes (measured on the ripa scene: 7.5 GB of DRAM reads in one 2.4 ms generation). The caller-facing ray batch
// stays SoA (optb_rays): it is read once, in order, fully coalesced.
struct __align__(16) RayRec {
  double ox, oy, oz, dx, dy, dz, I, wl, qre, qim, pl, n, len;
  uint32_t flags, root, pop;  // pop: pop number of the first ray of this root's generation (the ray's own = pop + rank)
  int32_t family;
  uint32_t rank, gcount;      // rank of the ray inside its root's generation / size of that generation (rank_kernel)
};
static_assert(sizeof(RayRec) == 128, "RayRec must be four sectors");
struct RayBuf {
  RayRec* rec;
  // Dense side arrays, one entry per record slot. The wavefront bookkeeping between two generations (mark, rank,
  // sort_prep) reads and writes ONLY these: 8 bytes per slot, visited in slot order, instead of one 32-byte sector of
  // every 128-byte record (ripa: 23.8 -> 21.7 ms per 1e6 rays).
  uint2* key;  // written with the child: x = coherence key (leaf it left) * 2 + child index, y = root
  uint2* rk;   // written by rank_kernel: x = rank of the ray inside its root's generation, y = size of that generation
};

struct Header {  // first bytes of the workspace
  unsigned int work_ctr;
  unsigned int n_next;      // size of the wavefront the last compaction produced
  unsigned int n_prev;      // size of the wavefront that compaction consumed (for generations launched unseen)
  unsigned int empty_gens;  // generations launched unseen that found nothing to do
  unsigned int pad[12];
};

struct SceneOff { uint32_t trav, nf, ni, matk, matf, mon, aux; };

struct TraceArgs {
  const unsigned char* blob; uint32_t blob_bytes; SceneOff off; int n_nodes, n_mons;
  // input wavefront: either the caller's rays (gen0) or the workspace wavefront
  optb_rays in0; RayBuf w; int gen0;
  long long n_in; const unsigned int* n_in_dev;  // n_in_dev overrides when non-null
  const uint32_t* gen_first; const uint32_t* gen_last;
  const uint32_t* perm;  // processing order of the wavefront (sorted by coherence key), or null
  const uint32_t* slot;  // wavefront index (reference order, dense) -> slot of the ray in the sparse buffer `w`
  // split output
  RayBuf c; uint8_t* nchild;
  // params
  long long max_trace; double unit; int rec_seg, rec_hit, rec_hist, chain_len, n_families, fam_shared;
  int hist_smem;
  optb_result out;
  unsigned long long* counters;
  Header* hdr;
  // family-serial mode (scenes with interact caps): one thread owns one Ray._id family and replays the
  // reference's sequential order; w is then a per-family FIFO ring of qcap entries
  const uint32_t* fam_off; uint32_t* fam_roots; uint32_t qcap;
  uint32_t root_base;  // global index of ray 0 of this launch (chunked host traces)
  int has_boxes;       // some node carries a lab-box test
};

OPTB_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

OPTB_DEV void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
OPTB_DEV void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
OPTB_DEV void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
OPTB_DEV void mbar_wait(unsigned long long* bar, uint32_t phase) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
      ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}

OPTB_DEV unsigned long long warp_alloc(unsigned long long* ctr) {
  unsigned m = __activemask();
  int lane = threadIdx.x & 31;
  int leader = __ffs(m) - 1;
  unsigned long long base = 0;
  if (lane == leader) base = atomicAdd(ctr, (unsigned long long)__popc(m));
  base = __shfl_sync(m, base, leader);
  return base + __popc(m & ((1u << lane) - 1u));
}

OPTB_DEV void load_ray(const TraceArgs& a, long long i, Ray& r, bool& solo, uint32_t& gcount, uint32_t& pop_base) {
  if (a.gen0) {
    const optb_rays& s = a.in0;
    const uint32_t bc = s.broadcast;
#define OPTB_COL(ptr, bit) ((ptr)[((bc >> (bit)) & 1u) ? 0 : i])
    r.ox = OPTB_COL(s.ox, 0); r.oy = OPTB_COL(s.oy, 1); r.oz = OPTB_COL(s.oz, 2);
    r.dx = OPTB_COL(s.dx, 3); r.dy = OPTB_COL(s.dy, 4); r.dz = OPTB_COL(s.dz, 5);
    r.I = OPTB_COL(s.intensity, 6); r.wl = OPTB_COL(s.wavelength, 7);
    r.qre = OPTB_COL(s.q_re, 8); r.qim = OPTB_COL(s.q_im, 9);
    r.pl = OPTB_COL(s.pathlength, 10); r.n = OPTB_COL(s.n_medium, 11);
    r.len = s.length ? OPTB_COL(s.length, 12) : INFINITY;
#undef OPTB_COL
    r.flags = s.flags ? s.flags[i] : (OPTB_RF_ALIVE | OPTB_RF_HASQ);
    r.root = (uint32_t)i + a.root_base; r.pop = 0;
    r.family = s.family ? s.family[i] : (int32_t)r.root;
    solo = true; gcount = 1; pop_base = 0;
  } else {
    // Children stay where the previous generation's kernel wrote them (slot 2*parent + k of the sparse buffer);
    // `slot` lists the occupied slots in reference order, so nothing is copied between generations.
    const uint32_t sl = a.slot[i];
    const uint2 rg = a.w.rk[sl];         // issued together with the record's loads (neither depends on the other)
    const RayRec q = a.w.rec[sl];        // eight 16-byte loads of one contiguous record
    r.ox = q.ox; r.oy = q.oy; r.oz = q.oz; r.dx = q.dx; r.dy = q.dy; r.dz = q.dz;
    r.I = q.I; r.wl = q.wl; r.qre = q.qre; r.qim = q.qim; r.pl = q.pl; r.n = q.n; r.len = q.len;
    r.flags = q.flags; r.root = q.root; r.family = q.family;
    gcount = rg.y;
    solo = (gcount == 1);
    pop_base = q.pop;
    r.pop = q.pop + rg.x;  // pop_base of this generation + rank inside the root
  }
}

OPTB_DEV void store_child(const RayBuf& c, long long j, const Ray& parent, double ox, double oy, double oz, double pl,
                           double dx, double dy, double dz, double I, double qre, double qim, double nmed, int k,
                           uint32_t pop_base, int leaf) {
  RayRec q;
  q.ox = ox; q.oy = oy; q.oz = oz; q.dx = dx; q.dy = dy; q.dz = dz;
  q.I = I; q.wl = parent.wl; q.qre = qre; q.qim = qim; q.pl = pl; q.n = nmed; q.len = parent.len;
  q.flags = parent.flags; q.root = parent.root; q.pop = pop_base; q.family = parent.family;
  q.rank = 0u; q.gcount = 1u;  // (unused by the wavefront: rank and generation size live in RayBuf::rk)
  c.rec[j] = q;                 // eight 16-byte stores: four full sectors
  c.key[j] = make_uint2(((uint32_t)leaf << 1) | (uint32_t)k, parent.root);
}

OPTB_DEV void ring_put(const RayBuf& w, long long j, const Ray& r) {
  RayRec q;
  q.ox = r.ox; q.oy = r.oy; q.oz = r.oz; q.dx = r.dx; q.dy = r.dy; q.dz = r.dz;
  q.I = r.I; q.wl = r.wl; q.qre = r.qre; q.qim = r.qim; q.pl = r.pl; q.n = r.n; q.len = r.len;
  q.flags = r.flags; q.root = r.root; q.pop = r.pop; q.family = r.family; q.rank = 0u; q.gcount = 1u;
  w.rec[j] = q;
}
OPTB_DEV void ring_get(const RayBuf& w, long long j, Ray& r) {
  const RayRec q = w.rec[j];
  r.ox = q.ox; r.oy = q.oy; r.oz = q.oz; r.dx = q.dx; r.dy = q.dy; r.dz = q.dz;
  r.I = q.I; r.wl = q.wl; r.qre = q.qre; r.qim = q.qim; r.pl = q.pl; r.n = q.n; r.len = q.len;
  r.flags = q.flags;
}

// One pop's dead segment: append to the segment log and test it against every monitor (monitor.py:183-193).
OPTB_COLD void emit_segment(const TraceArgs& a, const SceneView& sv, const Ray& r, double seg_len, uint32_t seg_flags,
                           int leaf, unsigned int* s_hist, unsigned int& n_hits_local) {
  if (a.rec_seg) {
    unsigned long long j = warp_alloc(&a.counters[OPTB_C_SEGMENTS]);
    if (j < (unsigned long long)a.out.seg_capacity) {
      const optb_result& o = a.out;
      o.seg_ox[j] = r.ox; o.seg_oy[j] = r.oy; o.seg_oz[j] = r.oz;
      o.seg_dx[j] = r.dx; o.seg_dy[j] = r.dy; o.seg_dz[j] = r.dz;
      o.seg_length[j] = seg_len; o.seg_intensity[j] = r.I; o.seg_wavelength[j] = r.wl;
      o.seg_q_re[j] = r.qre; o.seg_q_im[j] = r.qim; o.seg_pathlength[j] = r.pl; o.seg_n[j] = r.n;
      o.seg_flags[j] = seg_flags; o.seg_root[j] = r.root; o.seg_pop[j] = r.pop; o.seg_leaf[j] = leaf;
    } else {
      atomicOr(&a.counters[OPTB_C_STATUS], (unsigned long long)OPTB_ST_SEG_OVERFLOW);
    }
  }
  for (int m = 0; m < sv.n_mons; m++) {
    const double* mf = sv.mon + m * kMonBlobStride;
    // Monitor.record (monitor.py:183-193) = to-local + planar hit with t <= length, in stages like intersect_planar:
    // the x row of Tinv decides "in front, not parallel" and (orthonormal frame) gives t; most segments end here
    const double* c = mf + OPTB_MON_ORIGIN;
    const double* Ti = mf + OPTB_MON_TINV;
    const double vx = r.ox - c[0], vy = r.oy - c[1], vz = r.oz - c[2];
    const double ox = dot3(Ti[0], Ti[1], Ti[2], vx, vy, vz);
    double dx = dot3(Ti[0], Ti[1], Ti[2], r.dx, r.dy, r.dz);
    if (!((ox < 0.0 && dx > 0.0) || (ox > 0.0 && dx < 0.0))) continue;  // t = -ox/dx >= 1e-9 needs opposite signs
    double dy, dz;
    const bool ortho = mf[OPTB_MON_ORTHO] != 0.0;
    if (!ortho) {
      dy = dot3(Ti[3], Ti[4], Ti[5], r.dx, r.dy, r.dz);
      dz = dot3(Ti[6], Ti[7], Ti[8], r.dx, r.dy, r.dz);
      const double rn = rsqrt(dot3(dx, dy, dz, dx, dy, dz));
      dx *= rn; dy *= rn; dz *= rn;
    }
    const double t = -ox / dx;
    if (!(t >= 1e-9) || t > seg_len) continue;
    if (ortho) {
      dy = dot3(Ti[3], Ti[4], Ti[5], r.dx, r.dy, r.dz);
      dz = dot3(Ti[6], Ti[7], Ti[8], r.dx, r.dy, r.dz);
    }
    const double oy = dot3(Ti[3], Ti[4], Ti[5], vx, vy, vz);
    const double oz = dot3(Ti[6], Ti[7], Ti[8], vx, vy, vz);
    const double Px = fma(t, dx, ox), Py = fma(t, dy, oy), Pz = fma(t, dz, oz);
    if (!(fabs(Py) <= mf[OPTB_MON_HW] && fabs(Pz) <= mf[OPTB_MON_HH])) continue;
    if (a.rec_hist) {
      double y = dot3(Px, Py, Pz, mf[OPTB_MON_TY], mf[OPTB_MON_TY + 1], mf[OPTB_MON_TY + 2]);
      double z = dot3(Px, Py, Pz, mf[OPTB_MON_TZ], mf[OPTB_MON_TZ + 1], mf[OPTB_MON_TZ + 2]);
      int by = hist_bin(y, -mf[OPTB_MON_HW], mf[OPTB_MON_HW], mf[OPTB_MON_STRIDE], mf[OPTB_MON_STRIDE + 1]);
      int bz = hist_bin(z, -mf[OPTB_MON_HH], mf[OPTB_MON_HH], mf[OPTB_MON_STRIDE + 2], mf[OPTB_MON_STRIDE + 3]);
      const int per = OPTB_HIST_BINS * (OPTB_HIST_BINS + 1);
      if (by >= 0) {
        if (a.hist_smem) {
          atomicAdd(&s_hist[m * per + by], 1u);
          if (bz >= 0) atomicAdd(&s_hist[m * per + OPTB_HIST_BINS + by * OPTB_HIST_BINS + bz], 1u);
        } else {
          atomicAdd((unsigned long long*)&a.out.hist_y[m * OPTB_HIST_BINS + by], 1ull);
          if (bz >= 0)
            atomicAdd((unsigned long long*)&a.out.hist_yz[(m * OPTB_HIST_BINS + by) * OPTB_HIST_BINS + bz], 1ull);
        }
      }
    }
    if (a.rec_hit) {
      unsigned long long j = warp_alloc(&a.counters[OPTB_C_HITS]);
      if (j < (unsigned long long)a.out.hit_capacity) {
        const optb_result& o = a.out;
        if (o.hit_monitor) o.hit_monitor[j] = m;
        if (o.hit_key) o.hit_key[j] = ((unsigned long long)r.root << 32) | ((unsigned long long)m << 24) | (unsigned long long)r.pop;
        if (o.hit_root) o.hit_root[j] = r.root;
        if (o.hit_pop) o.hit_pop[j] = r.pop;
        if (o.hit_px) o.hit_px[j] = Px;
        if (o.hit_py) o.hit_py[j] = Py;
        if (o.hit_pz) o.hit_pz[j] = Pz;
        if (o.hit_intensity) o.hit_intensity[j] = r.I;
        if (o.hit_t) o.hit_t[j] = t;
        if (o.hit_dx) o.hit_dx[j] = r.dx;
        if (o.hit_dy) o.hit_dy[j] = r.dy;
        if (o.hit_dz) o.hit_dz[j] = r.dz;
        if (o.hit_q_re) o.hit_q_re[j] = r.qre;
        if (o.hit_q_im) o.hit_q_im[j] = r.qim;
      } else {
        atomicOr(&a.counters[OPTB_C_STATUS], (unsigned long long)OPTB_ST_HIT_OVERFLOW);
      }
    } else {
      n_hits_local++;
    }
  }
}

// Closest hit over the flattened tree (optical_table.py:119-123 + component_group.py:93-122): smallest t, ties to
// the smallest pre-order index. Leaves are visited in pre-order except that aspheres (a 10-sample scan plus an
// iterative root solve each) are parked and tested last, when the closest cheap hit is known: a parked asphere
// whose bracket starts beyond that hit is dismissed without evaluating its profile once. The (t, index) ordering
// makes the result independent of the visiting order.
template <bool ASPH, bool BRENT = false>
struct HitSearch {
  const TraceArgs& a; const SceneView& sv; const Ray& ray; bool solo;
  double best_t; int best_node; unsigned int* cnt;  // cnt: this thread's {leaf tests, curved tests, box tests} in smem

  OPTB_DEV void test_leaf(int i, const int32_t* __restrict__ ni, const double* __restrict__ nf) {
    atomicAdd(cnt, 1u);  // shared-memory add without a result: one instruction, no register held across the search
    const int slot = ni[OPTB_NI_CAPSLOT];
    // a capped surface counts every geometric hit, closest or not (optical_component.py:359-362): no early exit
    const double t_beat = slot >= 0 ? INFINITY : best_t;
    double t;
    if (OPTB_STAGED_PLANAR && is_planar_kind(ni[OPTB_NI_GEOM])) {
      t = intersect_planar(sv, ni, nf, ray, t_beat);
    } else {
      if (!is_planar_kind(ni[OPTB_NI_GEOM])) atomicAdd(cnt + kBlock, 1u);
      double ox, oy, oz, dx, dy, dz;
      to_local(nf + OPTB_NF_ORIGIN, nf + OPTB_NF_TINV, ray, ni[OPTB_NI_ORTHO] != 0, ox, oy, oz, dx, dy, dz);
      t = intersect_leaf<ASPH, BRENT>(sv, ni, nf, ox, oy, oz, dx, dy, dz, ray.len, t_beat);
    }
    if (!(t >= 0.0)) return;
    if (slot >= 0) {  // should_interact / increase_interact_count :136-149
      int32_t* cnt = a.out.cap_counts + (long long)slot * a.n_families + ray.family;
      int old = atomicAdd(cnt, 1);
      if (!((double)old < nf[OPTB_NF_CAPMAX])) {
        atomicSub(cnt, 1);
        if (!solo || a.fam_shared) atomicOr(&a.counters[OPTB_C_STATUS], (unsigned long long)OPTB_ST_CAP_ORDER);
        return;
      }
    }
    if (t < best_t || (t == best_t && i < best_node)) { best_t = t; best_node = i; }
  }
};

// Lattice window of a ray in an OPTB_G_GRID group (descriptor layout: include/optb.h OPTB_GRID_*): index ranges
// [i0, i1] x [j0, j1] of the children whose lab box the ray can touch. False = no cheap window (ray nearly in the
// lattice plane, or more than 64 cells): the caller descends the group's box hierarchy instead.
OPTB_DEV bool grid_window(const double* __restrict__ gd, const Ray& r, int& i0, int& i1, int& j0, int& j1) {
  const double R = gd[OPTB_GRID_R];
  const double px = r.ox - gd[OPTB_GRID_C00], py = r.oy - gd[OPTB_GRID_C00 + 1], pz = r.oz - gd[OPTB_GRID_C00 + 2];
  const double w0 = dot3(px, py, pz, gd[OPTB_GRID_NHAT], gd[OPTB_GRID_NHAT + 1], gd[OPTB_GRID_NHAT + 2]);
  const double wd = dot3(r.dx, r.dy, r.dz, gd[OPTB_GRID_NHAT], gd[OPTB_GRID_NHAT + 1], gd[OPTB_GRID_NHAT + 2]);
  if (!(fabs(wd) >= 1e-3)) return false;
  // |w0 + t wd| <= R, t >= 0
  const double inv = 1.0 / wd;
  double ta = (-R - w0) * inv, tb = (R - w0) * inv;
  if (ta > tb) { const double x = ta; ta = tb; tb = x; }
  i0 = j0 = 0; i1 = j1 = -1;
  if (!(tb >= 0.0)) return true;  // the lattice slab lies behind the ray: no lattice child can pass its box test
  ta = fmax(ta, 0.0);
  const double a0 = dot3(px, py, pz, gd[OPTB_GRID_UD], gd[OPTB_GRID_UD + 1], gd[OPTB_GRID_UD + 2]);
  const double ad = dot3(r.dx, r.dy, r.dz, gd[OPTB_GRID_UD], gd[OPTB_GRID_UD + 1], gd[OPTB_GRID_UD + 2]);
  const double b0 = dot3(px, py, pz, gd[OPTB_GRID_VD], gd[OPTB_GRID_VD + 1], gd[OPTB_GRID_VD + 2]);
  const double bd = dot3(r.dx, r.dy, r.dz, gd[OPTB_GRID_VD], gd[OPTB_GRID_VD + 1], gd[OPTB_GRID_VD + 2]);
  const double a1 = fma(ta, ad, a0), a2 = fma(tb, ad, a0), b1 = fma(ta, bd, b0), b2 = fma(tb, bd, b0);
  const double no = gd[OPTB_GRID_NOUTER] - 1.0, ni = gd[OPTB_GRID_NINNER] - 1.0;
  const double alo = ceil(fmin(a1, a2) - gd[OPTB_GRID_RHOA]), ahi = floor(fmax(a1, a2) + gd[OPTB_GRID_RHOA]);
  const double blo = ceil(fmin(b1, b2) - gd[OPTB_GRID_RHOB]), bhi = floor(fmax(b1, b2) + gd[OPTB_GRID_RHOB]);
  if (!(alo <= ahi && blo <= bhi)) return alo == alo && ahi == ahi && blo == blo && bhi == bhi;  // empty (NaN: no window)
  if (ahi < 0.0 || alo > no || bhi < 0.0 || blo > ni) return true;  // window beside the lattice
  i0 = (int)fmax(alo, 0.0); i1 = (int)fmin(ahi, no);
  j0 = (int)fmax(blo, 0.0); j1 = (int)fmin(bhi, ni);
  if ((i1 - i0 + 1) * (j1 - j0 + 1) > 64) { i1 = j1 = -1; return false; }
  return true;
}

// BOXES: 0 = the scene has no lab-box test at all (top-level leaves only);
//        1 = pre-order walk, a failed box test skips the subtree (component_group.py:98-115 as a forward scan);
//        2 = the same walk, and groups whose children form a regular lattice (OPTB_G_GRID: MMA / MLA / DMD arrays)
//            hand the walk the few children inside the ray's lattice window instead of a descent through their
//            box hierarchy; each candidate still has to pass the reference's own test of its own box.
template <int BOXES, bool ASPH, bool BRENT = false, bool VOTE = true>
OPTB_DEV void closest_hit(const TraceArgs& a, const SceneView& sv, const Ray& ray, bool solo,
                          double& best_t, int& best_node, unsigned int* cnt) {
  best_t = INFINITY; best_node = -1;
  const bool no_work = !(ray.flags & OPTB_RF_ALIVE);  // optical_component.py:349-350: a dead ray hits nothing
  HitSearch<ASPH, BRENT> hs{a, sv, ray, solo, INFINITY, -1, cnt};
  if constexpr (BOXES == 0 && !ASPH) {
    // No node carries a box test, so there is no group (groups always do): every node is a top-level leaf, and without
    // aspheres nothing is parked. The walk is a plain loop (c4: 687 -> ~600 warp instructions per pop).
    if (!no_work) {
      const int n = sv.n_nodes;
      // (one counter add per pop instead of one per leaf measured SLOWER on c4, 110.7 -> 113.6 ms: code layout)
#pragma unroll 1
      for (int i = 0; i < n; i++) hs.test_leaf(i, sv.ni + i * OPTB_NI_STRIDE, sv.nf + i * OPTB_NF_STRIDE);
    }
    best_t = hs.best_t; best_node = hs.best_node;
    return;
  }
  // front-to-back dismissal (slab_hit_far): 1 = in the walk and before a parked test; 2 = asphere variants only before a
  // parked test (their walks stay as they were)
  constexpr bool kCullWalk = OPTB_CULL_FAR == 1 || (OPTB_CULL_FAR == 2 && !ASPH);
  constexpr bool kCullPark = OPTB_CULL_FAR != 0;
  unsigned int n_box = 0;
  constexpr int kPark = 4;
  int parked[kPark];
  int n_parked = 0, n_done = 0;
  int i = 0;
  const int n = sv.n_nodes;
  // BOXES = 0: a scene of top-level leaves only has no box test at all (and pays no reciprocals for one)
  const BoxRay br(ray.ox, ray.oy, ray.oz, BOXES ? ray.dx : 1.0, BOXES ? ray.dy : 1.0, BOXES ? ray.dz : 1.0);
  // lattice cursor (BOXES = 2): group being listed (-1: none), cell cursor, window, cursor over the extra children
  int g_node = -1, g_i = 0, g_j = 0, g_i1 = -1, g_j0 = 0, g_j1 = -1, g_x = 0;
  unsigned long long g_mask = 0ull;  // (OPTB_WIN_BATCH) window cells whose box test passed and that are still to be visited
  const double* gd = nullptr;
  // Warp-synchronous "walk, then test": in every round each lane walks boxes until it holds a leaf (or runs out of
  // nodes and takes a parked asphere), then the lanes that called in together meet again at the vote and run the
  // one, long leaf test side by side. Without the explicit vote the compiler is free to fold the test into the
  // walk loop, and lanes that reach their leaves at different box counts then execute it one after another
  // (measured: 3x on the 7,689-leaf scene). One test_leaf call site also keeps the kernel's code size down.
  const unsigned group = __activemask();
  bool searching = !no_work;
  // (scenes without box tests yield a leaf per walk step: all lanes are in lock-step anyway, no vote needed)
  // VOTE = false (scenes staged in shared memory: a few dozen nodes, mostly coherent bundles): no vote either
  while ((BOXES && VOTE) ? __any_sync(group, searching) : searching) {
    int leaf = -1;
    if (searching) {
      while (true) {
        if (BOXES == 2 && g_node >= 0) {
          // next candidate child of the lattice group: window cells in list order, then the off-lattice children
          int cand;
          if (OPTB_WIN_BATCH && g_mask) {
            // (batched window: the boxes of all window cells were tested when the window was set up)
            const int b = __ffsll((long long)g_mask) - 1;
            g_mask &= g_mask - 1ull;
            cand = (int)gd[OPTB_GRID_CELLS + (g_i + (b >> 3)) * (int)gd[OPTB_GRID_NINNER] + g_j0 + (b & 7)];
            if (kCullWalk && (*reinterpret_cast<const int*>(sv.trav + cand * 8 + 7) & 2) && hs.best_t < INFINITY &&
                slab_far_only(br, sv.trav + cand * 8, hs.best_t)) continue;
            leaf = cand;
            break;
          }
          if (g_i <= g_i1) {
            cand = (int)gd[OPTB_GRID_CELLS + g_i * (int)gd[OPTB_GRID_NINNER] + g_j];
            if (++g_j > g_j1) { g_j = g_j0; g_i++; }
          } else if (g_x < (int)gd[OPTB_GRID_NEXT]) {
            cand = (int)gd[OPTB_GRID_CELLS + (int)gd[OPTB_GRID_NOUTER] * (int)gd[OPTB_GRID_NINNER] + g_x];
            g_x++;
          } else {  // done: leave the group's subtree
            i = reinterpret_cast<const int2*>(sv.trav + g_node * 8 + 6)->y;
            g_node = -1;
            continue;
          }
          n_box++;
          if (kCullWalk && (*reinterpret_cast<const int*>(sv.trav + cand * 8 + 7) & 2)) {
            if (slab_hit_far(br, sv.trav + cand * 8, hs.best_t) != 1) continue;
          } else if (!slab_hit(br, sv.trav + cand * 8)) continue;  // the child's own lab box (component_group.py:104-107)
          leaf = cand;  // (lattice children are leaves without interact caps, never aspheres to park: flatten._grid)
          break;
        }
        if (i >= n) break;
        const double* tv = sv.trav + i * 8;
        const int2 gs = *reinterpret_cast<const int2*>(tv + 6);  // {geometry kind, skip}
        if (BOXES) {
          const int word = *reinterpret_cast<const int*>(tv + 7);  // bit 0: test the box; bit 1: may dismiss by distance
          if (word) {
            n_box++;
            if (kCullWalk && (word & 2)) {
              // front to back: a cullable box entered beyond the closest hit so far cannot hold the winner
              if (slab_hit_far(br, tv, hs.best_t) != 1) { i = gs.y; continue; }
            } else if (!slab_hit(br, tv)) { i = gs.y; continue; }
          }
        }
        const int g = gs.x;
        const int cur = i++;
        if (BOXES == 2 && g == OPTB_G_GRID && !br.any_par) {
          gd = sv.aux + sv.ni[cur * OPTB_NI_STRIDE + OPTB_NI_AUX];
          int i0;
          if (grid_window(gd, ray, i0, g_i1, g_j0, g_j1)) {
            g_node = cur; g_i = i0; g_j = g_j0; g_x = 0;
            if (OPTB_WIN_BATCH && g_i1 - i0 < 8 && g_j1 - g_j0 < 8 && g_i1 >= i0 && g_j1 >= g_j0) {
              // The window's box tests as one batch, two at a time: they are independent (each child's own stored box,
              // component_group.py:104-107), and taken one by one each is a dependent chain of L2 round trips (cell
              // index -> box -> test -> branch). Bit (row * 8 + column) of g_mask = that cell's box is hit; the walk
              // then visits the set bits in list order and applies the front-to-back dismissal when it gets to each.
              const int nin = (int)gd[OPTB_GRID_NINNER];
              const int total = (g_i1 - i0 + 1) * (g_j1 - g_j0 + 1);
              int ii = i0, jj = g_j0;
              unsigned long long m = 0ull;
#pragma unroll 1
              for (int c = 0; c < total; c += 2) {
                int ii2 = ii, jj2 = jj + 1;
                if (jj2 > g_j1) { jj2 = g_j0; ii2++; }
                const bool two = c + 1 < total;
                if (!two) { ii2 = ii; jj2 = jj; }
                const int ca = (int)gd[OPTB_GRID_CELLS + ii * nin + jj];
                const int cb = (int)gd[OPTB_GRID_CELLS + ii2 * nin + jj2];
                const bool ha = slab_hit(br, sv.trav + ca * 8);
                const bool hb = slab_hit(br, sv.trav + cb * 8);
                m |= (unsigned long long)(ha ? 1u : 0u) << ((ii - i0) * 8 + (jj - g_j0));
                m |= (unsigned long long)((hb && two) ? 1u : 0u) << ((ii2 - i0) * 8 + (jj2 - g_j0));
                jj = jj2 + 1; ii = ii2;
                if (jj > g_j1) { jj = g_j0; ii++; }
              }
              n_box += (unsigned int)total;
              g_mask = m;
              g_i1 = i0 - 1;  // the cell cursor below is spent: after the mask come the off-lattice children
            }
          }
          continue;  // (no window: fall through into the box hierarchy below this node)
        }
        if (g == OPTB_G_GROUP || g == OPTB_G_GRID) continue;
        if (ASPH && g == OPTB_G_ASPHERE && n_parked < kPark) {
#pragma unroll
          for (int k = 0; k < kPark; k++) if (k == n_parked) parked[k] = cur;
          n_parked++;
          continue;
        }
        leaf = cur;
        break;
      }
      if (leaf < 0) {
        if (n_done < n_parked) {
#pragma unroll
          for (int k = 0; k < kPark; k++) if (k == n_done) leaf = parked[k];
          n_done++;
          // parked while the closest cheap hit was not known yet: its box may be out of reach by now
          if (kCullPark && BOXES && (*reinterpret_cast<const int*>(sv.trav + leaf * 8 + 7) & 2) &&
              slab_hit_far(br, sv.trav + leaf * 8, hs.best_t) == 2) leaf = -1;
        } else {
          searching = false;
        }
      }
    }
    if (leaf >= 0) hs.test_leaf(leaf, sv.ni + leaf * OPTB_NI_STRIDE, sv.nf + leaf * OPTB_NF_STRIDE);
  }
  best_t = hs.best_t; best_node = hs.best_node;
  if (BOXES) atomicAdd(cnt + 2 * kBlock, n_box);
}

// SPLIT = some interaction of the scene can emit two rays (or the caller bounded the in-register chain): the
// wavefront machinery (child slots, per-root generation ranks) is compiled in. Scenes that cannot split run the
// whole life of a ray in registers with none of it.
// FLAG = diagnostics variant: the ambiguity mask of SURVEY A.9 is evaluated at every pop (optb_flags.cuh).
// BRENT = params.reference_roots: curved-surface roots from the reference's own brentq iteration (brentq_dev) instead
// of the closed-form / Newton root.
// PASSK = the scene has an OPTB_I_PASS leaf (a Monitor listed as a component): only the general variants carry that body.
template <bool SMEM, bool SERIAL, int BOXES, bool ASPH, bool SPLIT, bool FLAG = false, bool BRENT = false, bool PASSK = false>
// Resident CTAs per SM: scenes staged in shared memory are bound by dependent fp64 chains and lose more to the
// spills of a tighter register budget than they gain from a fifth CTA (measured 6.32 -> 7.62 ms on c2); scenes
// whose tables stay in L1/L2 (thousands of leaves) are memory-latency bound and gain from it (ripa 30.6 -> 29.8 ms).
__global__ void __launch_bounds__(kBlock, kResident((SMEM || ASPH || SERIAL) ? OPTB_MIN_BLOCKS : OPTB_MIN_BLOCKS + OPTB_BIG_EXTRA))
trace_kernel(const __grid_constant__ TraceArgs a) {
  constexpr int MAXCH = (SPLIT || SERIAL) ? 2 : 1;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long mbar;
  // The closest-hit search only needs a ray's geometry. Its radiometric state (intensity, wavelength, q, path
  // length, index: 12 registers) is parked here for the duration of the search so that the hot loop has them.
  __shared__ double s_park[6][kBlock];
  __shared__ double s_icn[3][kBlock];   // index memo: wavelength, n of slot 0, n of slot 1
  __shared__ int s_icm[2][kBlock];      // material index of slot 0 / slot 1
  __shared__ unsigned int s_cnt[3][kBlock];  // per thread: leaf tests, curved leaf tests, lab-box tests
  s_cnt[0][threadIdx.x] = 0u; s_cnt[1][threadIdx.x] = 0u; s_cnt[2][threadIdx.x] = 0u;
  unsigned int* const my_cnt = &s_cnt[0][threadIdx.x];
  s_icn[0][threadIdx.x] = -1.0; s_icm[0][threadIdx.x] = -1; s_icm[1][threadIdx.x] = -1;
  const IndexCache ic{&s_icn[0][threadIdx.x], &s_icn[1][threadIdx.x], &s_icn[2][threadIdx.x], &s_icm[0][threadIdx.x],
                      &s_icm[1][threadIdx.x]};
  const unsigned char* base = a.blob;
  if constexpr (SMEM) {
    // Stage the whole scene blob with TMA bulk copies; completion is signalled on the mbarrier.
    if (threadIdx.x == 0) mbar_init(&mbar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
      mbar_expect_tx(&mbar, a.blob_bytes);
      constexpr uint32_t kChunk = 32768;
      for (uint32_t off = 0; off < a.blob_bytes; off += kChunk)
        tma_bulk_g2s(smem_raw + off, a.blob + off, min(kChunk, a.blob_bytes - off), &mbar);
    }
    mbar_wait(&mbar, 0);
    base = smem_raw;
  }
  SceneView sv;
  sv.trav = (const double*)(base + a.off.trav);
  sv.nf = (const double*)(base + a.off.nf);
  sv.ni = (const int32_t*)(base + a.off.ni);
  sv.matk = (const int32_t*)(base + a.off.matk);
  sv.matf = (const double*)(base + a.off.matf);
  sv.mon = (const double*)(base + a.off.mon);
  sv.aux = (const double*)(base + a.off.aux);
  sv.n_nodes = a.n_nodes; sv.n_mons = a.n_mons;
  sv.status = &a.counters[OPTB_C_STATUS];

  const int per = OPTB_HIST_BINS * (OPTB_HIST_BINS + 1);
  unsigned int* s_hist = (unsigned int*)(smem_raw + (SMEM ? ((a.blob_bytes + 15u) & ~15u) : 0u));
  if (a.hist_smem) {
    for (int k = threadIdx.x; k < a.n_mons * per; k += blockDim.x) s_hist[k] = 0u;
    __syncthreads();
  }

  const long long n_in = a.n_in_dev ? (long long)*a.n_in_dev : a.n_in;
  const int lane = threadIdx.x & 31;
  unsigned int c_pops = 0, c_inter = 0, c_drop = 0, c_hits = 0;  // per thread; widened when reduced

  while (true) {
    constexpr bool PHASE = OPTB_PHASE_SYNC && !SERIAL && !FLAG && BOXES != 0 && !ASPH;
    // REFILL: a lane whose ray is finished (its chain ended, or it was a one-pop ray) takes the next ray from the work
    // counter at once instead of idling until the slowest of its warp's 32 rays is done. Chains of a splitting scene end
    // after very different numbers of pops (ripa generation 0: 23.8 of 32 lanes busy on average). The work counter of
    // these variants counts rays, not 32-ray chunks.
    constexpr bool REFILL = OPTB_LANE_REFILL && PHASE && SPLIT;
    unsigned int chunk = 0;
    if (lane == 0) chunk = atomicAdd(&a.hdr->work_ctr, REFILL ? 32u : 1u);
    chunk = __shfl_sync(0xffffffffu, chunk, 0);
    long long i = (REFILL ? (long long)chunk : (long long)chunk * 32) + lane;
    // PHASE: the warps of a CTA start every pop together (one barrier per pop). The kernel's hot code (35-55 KB) is
    // larger than the 32 KB instruction cache; rays of one launch follow similar paths, so warps that start a pop
    // together run the same stretch of code at the same time and share the cache instead of thrashing it (ncu on the
    // ripa scene: 4.8 issue slots lost to instruction fetch per instruction issued). Used where it measured faster
    // (B200, r2n): scenes with a box walk and no aspheres -- lens groups of spheres (c3 -5 %), large arrays
    // (ripa -6 %); not for bare mirror scenes, whose pops are too short to pay for a barrier (c4 +17 %), nor for
    // asphere scenes (c2 +-0). A warp that ran out of rays keeps meeting the barrier until the whole CTA is done.
    const bool has = i < n_in;
    if (PHASE) { if (!__syncthreads_or(has ? 1 : 0)) break; }
    else {
      if ((long long)chunk * 32 >= n_in) break;
      if (!has) continue;
    }
    // Wavefront entries are stored in reference (BFS) order, which the pop numbering needs, but processed in the
    // order of their coherence key: rays that left the same surface the same way sit in the same warp.
    if (SPLIT && !SERIAL && a.perm && has) i = a.perm[i];

    if constexpr (SERIAL) {
      // Work item = one Ray._id family. Its initial rays are traced one after another in input order, each
      // with the reference's own FIFO loop (optical_table.py:74-147), so interact counts evolve exactly as
      // in the reference (optical_component.py:136-149, 359-362).
      const uint32_t lo = a.fam_off[i], hi = a.fam_off[i + 1];
      for (uint32_t x = lo + 1; x < hi; x++) {  // the scatter is unordered: restore input order
        uint32_t v = a.fam_roots[x]; uint32_t y = x;
        while (y > lo && a.fam_roots[y - 1] > v) { a.fam_roots[y] = a.fam_roots[y - 1]; y--; }
        a.fam_roots[y] = v;
      }
      const long long ring = (long long)i * a.qcap;
      for (uint32_t x = lo; x < hi; x++) {
        Ray ray; bool solo; uint32_t gcount, pop_base0;
        const uint32_t root = a.fam_roots[x];
        load_ray(a, root, ray, solo, gcount, pop_base0);
        uint32_t head = 0, tail = 0, pops = 0;
        ring_put(a.w, ring, ray); tail = 1;
        while ((long long)pops < a.max_trace && head != tail) {
          ring_get(a.w, ring + head % a.qcap, ray); head++;
          ray.root = root; ray.family = (int32_t)i; ray.pop = pops++;
          double t; int node;
          closest_hit<BOXES, ASPH, BRENT>(a, sv, ray, true, t, node, my_cnt);
          c_pops++;
          const bool hit = node >= 0;
          const int32_t* ni = sv.ni + (hit ? node : 0) * OPTB_NI_STRIDE;
          const double* nf = sv.nf + (hit ? node : 0) * OPTB_NF_STRIDE;
          emit_segment(a, sv, ray, hit ? t : ray.len, hit ? (ray.flags & ~OPTB_RF_ALIVE) : ray.flags,
                       hit ? ni[OPTB_NI_LEAF] : -1, s_hist, c_hits);
          if (!hit) continue;
          c_inter++;
          double ox, oy, oz, dx, dy, dz;
          to_local(nf + OPTB_NF_ORIGIN, nf + OPTB_NF_TINV, ray, ni[OPTB_NI_ORTHO] != 0, ox, oy, oz, dx, dy, dz);
          Children<MAXCH> ch;
          interact<ASPH, MAXCH, PASSK || SERIAL || FLAG || BRENT>(sv, ni, nf, ray, a.unit, ox, oy, oz, dx, dy, dz, t, ch, ic);
          for (int k = 0; k < ch.n; k++) {
            if (tail - head >= a.qcap) { atomicOr(&a.counters[OPTB_C_STATUS], (unsigned long long)OPTB_ST_WORK_OVERFLOW); break; }
            Ray c = ray;
            c.ox = ch.ox; c.oy = ch.oy; c.oz = ch.oz;
            const bool second = (k == 1);
            c.dx = second ? ch.dx[MAXCH - 1] : ch.dx[0]; c.dy = second ? ch.dy[MAXCH - 1] : ch.dy[0];
            c.dz = second ? ch.dz[MAXCH - 1] : ch.dz[0]; c.I = second ? ch.I[MAXCH - 1] : ch.I[0];
            c.qre = second ? ch.qre[MAXCH - 1] : ch.qre[0]; c.qim = second ? ch.qim[MAXCH - 1] : ch.qim[0];
            c.n = second ? ch.nmed[MAXCH - 1] : ch.nmed[0]; c.pl = ch.pl;
            ring_put(a.w, ring + tail % a.qcap, c); tail++;
          }
        }
        c_drop += tail - head;
      }
      continue;
    }

    Ray ray; bool solo = true; uint32_t gcount = 1, pop_base0 = 0;
    if (has) load_ray(a, i, ray, solo, gcount, pop_base0);
    if (!SPLIT) solo = true;
    uint32_t pop_base_next = (!SPLIT || a.gen0) ? 0u : (pop_base0 + gcount);
    int nch = 0;
    int chained = 0;
    int hit_leaf = 0;
    Children<MAXCH> ch;
    ch.n = 0;
    // one pop of this lane's ray; true = its only child carries on in registers (chain)
    auto pop_step = [&]() -> bool {
      if ((long long)ray.pop >= a.max_trace) { c_drop++; nch = 0; return false; }  // queued but never popped
      double t; int node;
      if constexpr (OPTB_PARK_ALWAYS || ASPH || BOXES == 2) {
        volatile double* pk = &s_park[0][threadIdx.x];
        pk[0 * kBlock] = ray.I; pk[1 * kBlock] = ray.wl; pk[2 * kBlock] = ray.qre;
        pk[3 * kBlock] = ray.qim; pk[4 * kBlock] = ray.pl; pk[5 * kBlock] = ray.n;
        closest_hit<BOXES, ASPH, BRENT, OPTB_VOTE_SMEM || SMEM != 1>(a, sv, ray, solo, t, node, my_cnt);
        ray.I = pk[0 * kBlock]; ray.wl = pk[1 * kBlock]; ray.qre = pk[2 * kBlock];
        ray.qim = pk[3 * kBlock]; ray.pl = pk[4 * kBlock]; ray.n = pk[5 * kBlock];
      } else {
        closest_hit<BOXES, ASPH, BRENT, OPTB_VOTE_SMEM || SMEM != 1>(a, sv, ray, solo, t, node, my_cnt);
      }
      c_pops++;
      if constexpr (FLAG) {
        const unsigned amb = flag_pop(sv, ray, node, t, a.unit);
        if (amb && a.out.root_flags) atomicOr(&a.out.root_flags[ray.root - a.root_base], amb);
      }
      // the pop's dead segment: the ray itself when nothing was hit (optical_table.py:132-134), else the
      // truncated copy with length = t, alive = False (optical_component.py:364)
      const bool hit = node >= 0;
      const int32_t* ni = sv.ni + (hit ? node : 0) * OPTB_NI_STRIDE;
      const double* nf = sv.nf + (hit ? node : 0) * OPTB_NF_STRIDE;
      emit_segment(a, sv, ray, hit ? t : ray.len, hit ? (ray.flags & ~OPTB_RF_ALIVE) : ray.flags,
                   hit ? ni[OPTB_NI_LEAF] : -1, s_hist, c_hits);
      if (!hit) { nch = 0; return false; }
      c_inter++;
      hit_leaf = ni[OPTB_NI_LEAF];
      double ox, oy, oz, dx, dy, dz;
      to_local(nf + OPTB_NF_ORIGIN, nf + OPTB_NF_TINV, ray, ni[OPTB_NI_ORTHO] != 0, ox, oy, oz, dx, dy, dz);
      interact<ASPH, MAXCH, PASSK || SERIAL || FLAG || BRENT>(sv, ni, nf, ray, a.unit, ox, oy, oz, dx, dy, dz, t, ch, ic);
      nch = ch.n;
      if (nch == 1 && solo && (!SPLIT || a.chain_len == 0 || chained + 1 < a.chain_len)) {
        // the root's alive set is this one ray: BFS order is trivially kept, continue in registers
        ray.ox = ch.ox; ray.oy = ch.oy; ray.oz = ch.oz;
        ray.dx = ch.dx[0]; ray.dy = ch.dy[0]; ray.dz = ch.dz[0];
        ray.I = ch.I[0]; ray.qre = ch.qre[0]; ray.qim = ch.qim[0]; ray.pl = ch.pl; ray.n = ch.nmed[0];
        ray.pop++; chained++;
        return true;
      }
      return false;
    };
    // what a finished ray leaves behind: its children in slots 2i / 2i + 1 of the sparse buffer
    auto store_children = [&]() {
      a.nchild[i] = (uint8_t)nch;
      uint32_t pb = solo ? ray.pop + 1u : pop_base_next;
      if (nch > 0) store_child(a.c, 2 * i, ray, ch.ox, ch.oy, ch.oz, ch.pl, ch.dx[0], ch.dy[0], ch.dz[0], ch.I[0],
                               ch.qre[0], ch.qim[0], ch.nmed[0], 0, pb, hit_leaf);
      if (nch > 1) store_child(a.c, 2 * i + 1, ray, ch.ox, ch.oy, ch.oz, ch.pl, ch.dx[MAXCH - 1], ch.dy[MAXCH - 1],
                               ch.dz[MAXCH - 1], ch.I[MAXCH - 1], ch.qre[MAXCH - 1], ch.qim[MAXCH - 1],
                               ch.nmed[MAXCH - 1], 1, pb, hit_leaf);
    };
    if constexpr (REFILL) {
      bool active = has;
      bool more = true;  // (warp-uniform) the work counter has not run past the wavefront yet
      while (__syncthreads_or(active ? 1 : 0)) {
        if (active) {
          active = pop_step();
          if (!active && a.nchild) store_children();
        }
        const unsigned need = more ? __ballot_sync(0xffffffffu, !active) : 0u;
        if (need) {
          const int leader = __ffs(need) - 1;
          unsigned int base = 0;
          if (lane == leader) base = atomicAdd(&a.hdr->work_ctr, (unsigned int)__popc(need));
          base = __shfl_sync(0xffffffffu, base, leader);
          if ((long long)base + __popc(need) >= n_in) more = false;
          if (!active) {
            i = (long long)base + __popc(need & ((1u << lane) - 1u));
            if (i < n_in) {
              if (a.perm) i = a.perm[i];
              load_ray(a, i, ray, solo, gcount, pop_base0);
              pop_base_next = a.gen0 ? 0u : (pop_base0 + gcount);
              nch = 0; chained = 0; hit_leaf = 0; ch.n = 0;
              active = true;
            }
          }
        }
      }
      continue;  // (the next trip finds the counter exhausted and leaves through the barrier above)
    } else if constexpr (PHASE) {
      bool active = has;
      while (__syncthreads_or(active ? 1 : 0)) {
        if (active) active = pop_step();
      }
    } else {
      while (pop_step()) {}
    }
    if constexpr (SPLIT) {
      if (a.nchild && has) {
        a.nchild[i] = (uint8_t)nch;
        uint32_t pb = solo ? ray.pop + 1u : pop_base_next;
        if (nch > 0) store_child(a.c, 2 * i, ray, ch.ox, ch.oy, ch.oz, ch.pl, ch.dx[0], ch.dy[0], ch.dz[0], ch.I[0],
                                 ch.qre[0], ch.qim[0], ch.nmed[0], 0, pb, hit_leaf);
        if (nch > 1) store_child(a.c, 2 * i + 1, ray, ch.ox, ch.oy, ch.oz, ch.pl, ch.dx[MAXCH - 1], ch.dy[MAXCH - 1],
                                 ch.dz[MAXCH - 1], ch.I[MAXCH - 1], ch.qre[MAXCH - 1], ch.qim[MAXCH - 1],
                                 ch.nmed[MAXCH - 1], 1, pb, hit_leaf);
      }
    }
  }

  // counters: warp reduce then one atomic per warp
  unsigned long long v[7] = {c_pops, c_inter, my_cnt[0], c_drop, c_hits, my_cnt[kBlock], my_cnt[2 * kBlock]};
#pragma unroll
  for (int k = 0; k < 7; k++) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], s);
  }
  if (lane == 0) {
    if (!a.rec_seg && v[0]) atomicAdd(&a.counters[OPTB_C_SEGMENTS], v[0]);
    if (v[1]) atomicAdd(&a.counters[OPTB_C_INTERACTIONS], v[1]);
    if (v[2]) atomicAdd(&a.counters[OPTB_C_TESTS], v[2]);
    if (v[3]) atomicAdd(&a.counters[OPTB_C_DROPPED], v[3]);
    if (!a.rec_hit && v[4]) atomicAdd(&a.counters[OPTB_C_HITS], v[4]);
    if (v[5]) atomicAdd(&a.counters[OPTB_C_TESTS_CURVED], v[5]);
    if (v[6]) atomicAdd(&a.counters[OPTB_C_BOX_TESTS], v[6]);
  }
  if (a.hist_smem) {
    __syncthreads();
    for (int k = threadIdx.x; k < a.n_mons * per; k += blockDim.x) {
      unsigned int cnt = s_hist[k];
      if (!cnt) continue;
      int m = k / per, r = k % per;
      if (r < OPTB_HIST_BINS) atomicAdd((unsigned long long*)&a.out.hist_y[m * OPTB_HIST_BINS + r], (unsigned long long)cnt);
      else atomicAdd((unsigned long long*)&a.out.hist_yz[m * OPTB_HIST_BINS * OPTB_HIST_BINS + (r - OPTB_HIST_BINS)],
                     (unsigned long long)cnt);
    }
  }
}

// ---- ordered compaction of the sparse children into the next wavefront --------------------------------
__global__ void __launch_bounds__(kScanBlock) tile_sums_kernel(const uint8_t* __restrict__ nchild, long long n,
                                                               unsigned int* __restrict__ sums, const Header* unseen) {
  __shared__ unsigned int s[kScanBlock / 32];
  if (unseen) n = unseen->n_next;  // generation launched without the host knowing its size
  long long base = (long long)blockIdx.x * kTile;
  unsigned int acc = 0;
  for (int k = threadIdx.x; k < kTile; k += kScanBlock) {
    long long i = base + k;
    if (i < n) acc += nchild[i];
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int t = 0;
    for (int w = 0; w < kScanBlock / 32; w++) t += s[w];
    sums[blockIdx.x] = t;
  }
}

// single block: exclusive scan of the tile sums in place; total -> hdr->n_next; also resets the work counter
__global__ void __launch_bounds__(1024) scan_sums_kernel(unsigned int* __restrict__ sums, int ntiles, Header* hdr,
                                                         unsigned long long* counters, unsigned int capacity,
                                                         bool unseen = false) {
  __shared__ unsigned int s[1024];
  __shared__ unsigned int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int b0 = 0; b0 < ntiles; b0 += 1024) {
    int idx = b0 + threadIdx.x;
    unsigned int v = idx < ntiles ? sums[idx] : 0u;
    s[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      unsigned int t = threadIdx.x >= o ? s[threadIdx.x - o] : 0u;
      __syncthreads();
      s[threadIdx.x] += t;
      __syncthreads();
    }
    unsigned int incl = s[threadIdx.x];
    if (idx < ntiles) sums[idx] = carry + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    unsigned int total = carry;
    if (total > capacity) {
      atomicOr(&counters[OPTB_C_STATUS], (unsigned long long)OPTB_ST_WORK_OVERFLOW);
      total = 0;  // stop cleanly
    }
    if (unseen) {
      hdr->n_prev = hdr->n_next;
      if (hdr->n_next == 0) hdr->empty_gens++;
    }
    hdr->n_next = total;
    hdr->work_ctr = 0;
  }
}

// Ordered compaction, index only: the k-th child of wavefront entry i sits in slot 2i+k of the sparse buffer; this
// writes the list of occupied slots in reference (BFS) order. Thread t of a block handles entries base + sub*256 + t.
__global__ void __launch_bounds__(kScanBlock) slots_kernel(const uint8_t* __restrict__ nchild, long long n,
                                                           const unsigned int* __restrict__ tile_base,
                                                           uint32_t* __restrict__ slot_next, const Header* hdr,
                                                           bool unseen) {
  __shared__ unsigned int s_warp[kScanBlock / 32];
  __shared__ unsigned int s_carry;
  if (hdr->n_next == 0) return;
  if (unseen) n = hdr->n_prev;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = tile_base[blockIdx.x];
  __syncthreads();
  for (int sub = 0; sub < kTile / kScanBlock; sub++) {
    const long long i = (long long)blockIdx.x * kTile + (long long)sub * kScanBlock + threadIdx.x;
    const unsigned int mine = (i < n) ? nchild[i] : 0u;
    unsigned int incl = mine;
    for (int o = 1; o < 32; o <<= 1) {
      unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    unsigned int woff = 0, total = 0;
    for (int k = 0; k < kScanBlock / 32; k++) { if (k < wid) woff += s_warp[k]; total += s_warp[k]; }
    const unsigned int off = s_carry + woff + incl - mine;
    for (unsigned int j = 0; j < mine; j++) slot_next[off + j] = (uint32_t)(2 * i + j);
    __syncthreads();
    if (threadIdx.x == 0) s_carry += total;
    __syncthreads();
  }
}

__global__ void mark_kernel(const uint2* __restrict__ key, const uint32_t* __restrict__ slot, const Header* hdr,
                            uint32_t* __restrict__ gen_first, uint32_t* __restrict__ gen_last) {
  long long n = hdr->n_next;
  for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) {
    uint32_t r = key[slot[j]].y;
    if (j == 0 || key[slot[j - 1]].y != r) gen_first[r] = (uint32_t)j;
    if (j == n - 1 || key[slot[j + 1]].y != r) gen_last[r] = (uint32_t)j;
  }
}
// rank of every wavefront entry inside its root's generation and the size of that generation, next to the ray's
// record (same slot): the trace kernel loads both at once, with no per-root lookup behind the record's root
__global__ void rank_kernel(const uint2* __restrict__ key, uint2* __restrict__ rk, const uint32_t* __restrict__ slot,
                            const Header* hdr, const uint32_t* __restrict__ gen_first, const uint32_t* __restrict__ gen_last) {
  long long n = hdr->n_next;
  for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) {
    const uint32_t s = slot[j];
    const uint32_t r = key[s].y, first = gen_first[r];
    rk[s] = make_uint2((uint32_t)j - first, gen_last[r] - first + 1u);
  }
}

// family -> list of initial rays (CSR) for the family-serial mode
__global__ void fam_count_kernel(const int32_t* __restrict__ family, long long n, unsigned int* __restrict__ counts) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    atomicAdd(&counts[family ? family[i] : (int32_t)i], 1u);
}
__global__ void fam_scatter_kernel(const int32_t* __restrict__ family, long long n, const unsigned int* __restrict__ off,
                                   unsigned int* __restrict__ cursor, uint32_t* __restrict__ roots) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int f = family ? family[i] : (int32_t)i;
    roots[off[f] + atomicAdd(&cursor[f], 1u)] = (uint32_t)i;
  }
}

// Sort input of one generation: identity permutation + the coherence key, with "this root has a single live ray"
// (it will chain many pops in registers) as the top bit so that one-pop rays and chaining rays do not share warps.
__global__ void sort_prep_kernel(uint32_t* __restrict__ idx, uint32_t* __restrict__ key_dense, const uint2* __restrict__ key,
                                 const uint2* __restrict__ rk, const uint32_t* __restrict__ slot, long long n, int key_bits) {
  for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) {
    idx[j] = (uint32_t)j;
    const uint32_t s = slot[j];
    const bool solo = rk[s].y == 1u;
    key_dense[j] = (key[s].x & ((1u << key_bits) - 1u)) | ((solo ? 1u : 0u) << key_bits);
  }
}

__global__ void count_flags_kernel(const uint32_t* __restrict__ flags, long long n, unsigned long long* counters) {
  unsigned int c = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) c += flags[i] != 0u;
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&counters[OPTB_C_FLAGGED], (unsigned long long)c);
}

__global__ void finish_kernel(unsigned long long* counters, unsigned long long gens, unsigned long long launches,
                              const Header* hdr) {
  counters[OPTB_C_GENERATIONS] = gens - hdr->empty_gens;
  counters[OPTB_C_LAUNCHES] = launches;
}

// FP64 FMA peak probe: 8 independent chains per thread
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters) {
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 0.999999, c = 1e-6;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace

// ---- host side ----------------------------------------------------------------------------------------
struct optb_ctx {
  int device;
  int sm_count;
  size_t smem_optin;
  char err[512];
  // arena for optb_trace_host
  void* arena; size_t arena_bytes;
  unsigned long long* h_counters;  // pinned
  unsigned int* h_hdr;             // pinned
  cudaStream_t s_h2d, s_run, s_d2h;  // pipelined optb_trace_host
  unsigned long long* h_chunk; size_t h_chunk_n;  // pinned per-chunk counters
  // Retired scene blobs, kept for the next upload: cudaMalloc/cudaFree cost milliseconds each (and cudaFree
  // synchronises the device), which is most of the latency of a small trace that re-uploads its scene per call.
  struct { unsigned char* p; size_t cap; } pool[8];
  // live-ray budget per initial ray that the last splitting optb_trace_host call needed (its grow loop starts there)
  int64_t live_per_ray_hint;
  // NCCL communicator for optb_monitor_merge (library loaded at run time)
  void* nccl_lib; void* nccl_comm; int comm_rank, comm_size; long long* d_merge;  // d_merge: device scratch of the merge
};

struct optb_scene {
  unsigned char* d_blob; size_t blob_cap; uint32_t blob_bytes; SceneOff off;
  int n_nodes, n_leaves, n_mats, n_mons, n_caps; long long n_aux;
  int max_children; int has_boxes; int has_asph; int has_grid; int has_pass;
  bool in_smem; bool hist_smem; uint32_t smem_bytes;
  // recorded on the stream of every trace that reads the blob: releasing the scene waits for this event only,
  // not for the whole device (other streams, NCCL and unrelated kernels keep running)
  cudaEvent_t last_use; bool used;
  std::vector<unsigned char>* cull;  // cullable bit per node as uploaded (optb_scene_update_nodes keeps it current)
};

static int fail(optb_ctx* ctx, int code, const char* what, cudaError_t e = cudaSuccess) {
  if (ctx) {
    if (e != cudaSuccess) snprintf(ctx->err, sizeof ctx->err, "%s: %s", what, cudaGetErrorString(e));
    else snprintf(ctx->err, sizeof ctx->err, "%s", what);
  }
  return code;
}
#define CK(call, what) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return fail(ctx, -10, what, e__); } while (0)

extern "C" int optb_abi_version(void) { return OPTB_ABI_VERSION; }
extern "C" int optb_comm_destroy(optb_ctx* ctx);
extern "C" int64_t optb_sort_workspace_bytes(int64_t n_rows);
static int sort_rows_dev(optb_ctx* ctx, const optb_result* res, const optb_result* dst, int64_t n_seg, int64_t n_hit,
                         const unsigned long long* d_counters, void* workspace, int64_t workspace_bytes, cudaStream_t st,
                         int end_bit);
extern "C" int optb_sort_rows(optb_ctx* ctx, const optb_result* res, int64_t n_seg, int64_t n_hit, void* workspace,
                              int64_t workspace_bytes, void* stream_v);

extern "C" int optb_ctx_create(int device, optb_ctx** out) {
  if (!out) return -1;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || device < 0 || device >= ndev) return -2;
  optb_ctx* ctx = new (std::nothrow) optb_ctx();
  if (!ctx) return -3;
  memset(ctx, 0, sizeof *ctx);
  ctx->device = device;
  int sm = 0, optin = 0;
  if (cudaSetDevice(device) != cudaSuccess ||
      cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, device) != cudaSuccess ||
      cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) != cudaSuccess ||
      cudaMallocHost((void**)&ctx->h_counters, sizeof(unsigned long long) * OPTB_C_COUNT) != cudaSuccess ||
      cudaMallocHost((void**)&ctx->h_hdr, sizeof(Header)) != cudaSuccess) {
    cudaGetLastError();
    optb_ctx_destroy(ctx);
    return -10;
  }
  ctx->sm_count = sm;
  ctx->smem_optin = (size_t)optin;
  *out = ctx;
  return 0;
}

extern "C" int optb_ctx_destroy(optb_ctx* ctx) {
  if (!ctx) return 0;
  cudaSetDevice(ctx->device);
  optb_comm_destroy(ctx);
  if (ctx->arena) cudaFree(ctx->arena);
  for (auto& b : ctx->pool) if (b.p) cudaFree(b.p);
  if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
  if (ctx->h_hdr) cudaFreeHost(ctx->h_hdr);
  if (ctx->h_chunk) cudaFreeHost(ctx->h_chunk);
  if (ctx->s_h2d) { cudaStreamDestroy(ctx->s_h2d); cudaStreamDestroy(ctx->s_run); cudaStreamDestroy(ctx->s_d2h); }
  delete ctx;
  return 0;
}

extern "C" const char* optb_last_error(const optb_ctx* ctx) { return ctx ? ctx->err : "null context"; }

// Which boxes may ALSO be used to dismiss a subtree by distance ("cullable"): a hit on a leaf always lies inside the
// leaf's geometric extent (planar: the aperture in the plane x = 0; curved: the local box that brackets the reference's
// root search, optical_component.py:197-233), so if the stored lab box contains that extent in the leaf's CURRENT pose,
// every hit on the leaf is at t >= the box's entry parameter, and a box entered beyond the closest hit found so far
// cannot change the result. The reference's boxes are cached and may be stale (the flattener ships the
// object's own `.bbox`): a stale box is still TESTED like the reference does, but never used for dismissal. A group is
// cullable when its box contains the boxes of all its children, all of them cullable. Capped leaves count every
// geometric hit, closest or not (optical_component.py:359-362): never cullable, nor is any group above them.
static void cull_bits(const optb_scene_desc* d, std::vector<unsigned char>& cull) {
  const int n = d->n_nodes;
  cull.assign(n, 0);
  auto contains = [](const double* outer, const double* inner) {
    for (int ax = 0; ax < 3; ax++) {
      const double tol = 1e-9 * std::max(1.0, std::max(fabs(inner[2 * ax]), fabs(inner[2 * ax + 1])));
      if (!(outer[2 * ax] <= inner[2 * ax] + tol && outer[2 * ax + 1] >= inner[2 * ax + 1] - tol)) return false;
    }
    return true;
  };
  for (int i = n - 1; i >= 0; i--) {  // children before parents (pre-order table, reverse scan)
    const int32_t* ni = d->node_i + (size_t)i * OPTB_NI_STRIDE;
    const double* nf = d->node_f + (size_t)i * OPTB_NF_STRIDE;
    if (!ni[OPTB_NI_AABB]) continue;
    const int g = ni[OPTB_NI_GEOM];
    if (g == OPTB_G_GROUP || g == OPTB_G_GRID) {
      bool ok = true;
      for (int j = i + 1; j < ni[OPTB_NI_SKIP] && ok; j = d->node_i[(size_t)j * OPTB_NI_STRIDE + OPTB_NI_SKIP])
        ok = cull[j] && contains(nf + OPTB_NF_AABB, d->node_f + (size_t)j * OPTB_NF_STRIDE + OPTB_NF_AABB);
      cull[i] = ok && ni[OPTB_NI_SKIP] > i + 1;
      continue;
    }
    if (ni[OPTB_NI_CAPSLOT] >= 0) continue;
    const double* p = nf + OPTB_NF_P;
    double e[6];  // geometric extent of the hittable points in the local frame
    switch (g) {
      case OPTB_G_CIRCLE: e[0] = e[1] = 0; e[2] = e[4] = -p[0]; e[3] = e[5] = p[0]; break;
      case OPTB_G_RECT: e[0] = e[1] = 0; e[2] = -p[0]; e[3] = p[0]; e[4] = -p[1]; e[5] = p[1]; break;
      case OPTB_G_SPHERE: for (int k = 0; k < 6; k++) e[k] = p[2 + k]; break;
      case OPTB_G_ASPHERE: e[0] = p[6]; e[1] = p[7]; e[2] = e[4] = -p[0]; e[3] = e[5] = p[0]; break;
      case OPTB_G_CYL: e[0] = e[2] = -p[0]; e[1] = e[3] = p[0]; e[4] = -p[1] / 2; e[5] = p[1] / 2; break;
      default: continue;  // polygons, composite apertures: boxes are tested, not used for dismissal
    }
    if (!(e[0] <= e[1] && e[2] <= e[3] && e[4] <= e[5])) continue;
    const double* T = nf + OPTB_NF_T;
    const double* c = nf + OPTB_NF_ORIGIN;
    double lab[6] = {INFINITY, -INFINITY, INFINITY, -INFINITY, INFINITY, -INFINITY};
    for (int corner = 0; corner < 8; corner++) {
      const double x = e[corner & 1], y = e[2 + ((corner >> 1) & 1)], z = e[4 + ((corner >> 2) & 1)];
      for (int ax = 0; ax < 3; ax++) {
        const double v = T[3 * ax] * x + T[3 * ax + 1] * y + T[3 * ax + 2] * z + c[ax];
        lab[2 * ax] = std::min(lab[2 * ax], v); lab[2 * ax + 1] = std::max(lab[2 * ax + 1], v);
      }
    }
    // (a curved leaf finds roots up to 1e-9 outside its bracket; the tolerance of `contains` covers that)
    cull[i] = contains(nf + OPTB_NF_AABB, lab);
  }
}

// Device-side extras of an asphere row: k1 of the sign / Newton functor (AsphF) in the FOCAL slot, which a refractive
// leaf does not use. Same expressions, same IEEE operations as the device used to evaluate per test.
static void asphere_constants(double* nf, const int32_t* ni) {
  if (ni[OPTB_NI_GEOM] != OPTB_G_ASPHERE) return;
  const double* c = nf + OPTB_NF_P + 1;
  if (ni[OPTB_NI_AUX] == OPTB_ASPH_PARAMETRIC) nf[OPTB_NF_FOCAL] = (1.0 + c[1]) / (c[0] * c[0]);
  else nf[OPTB_NF_FOCAL] = (c[1] + 1.0) / ((c[1] - 1.0) * c[0] * c[0]);
}

extern "C" int optb_scene_upload(optb_ctx* ctx, const optb_scene_desc* d, optb_scene** out) {
  if (!ctx || !d || !out) return -1;
  if (d->abi_version != OPTB_ABI_VERSION) return fail(ctx, -4, "scene ABI version mismatch");
  if (d->n_nodes < 0 || d->n_materials < 1) return fail(ctx, -5, "bad scene counts");
  // The kernels index the tables without bounds checks: reject anything that could walk out of them.
  for (int i = 0; i < d->n_nodes; i++) {
    const int32_t* ni = d->node_i + (size_t)i * OPTB_NI_STRIDE;
    const int g = ni[OPTB_NI_GEOM];
    if (ni[OPTB_NI_SKIP] <= i || ni[OPTB_NI_SKIP] > d->n_nodes) return fail(ctx, -5, "scene: node skip pointer out of range");
    if (g < OPTB_G_GROUP || g > OPTB_G_GRID) return fail(ctx, -5, "scene: unknown geometry kind");
    if (g == OPTB_G_GRID) {  // lattice descriptor: inside the aux pool, every listed child a leaf of this subtree
      const long long off = ni[OPTB_NI_AUX];
      if (off < 0 || off + OPTB_GRID_CELLS > d->n_aux) return fail(ctx, -5, "scene: lattice descriptor out of range");
      const double* gd = d->aux + off;
      const double cells = gd[OPTB_GRID_NOUTER] * gd[OPTB_GRID_NINNER] + gd[OPTB_GRID_NEXT];
      if (!(gd[OPTB_GRID_NOUTER] >= 1 && gd[OPTB_GRID_NINNER] >= 1 && gd[OPTB_GRID_NEXT] >= 0 && cells <= 1e8) ||
          off + OPTB_GRID_CELLS + (long long)cells > d->n_aux)
        return fail(ctx, -5, "scene: lattice descriptor out of range");
      for (long long k = 0; k < (long long)cells; k++) {
        const double c = gd[OPTB_GRID_CELLS + k];
        if (!(c > i && c < ni[OPTB_NI_SKIP]) || d->node_i[(size_t)c * OPTB_NI_STRIDE + OPTB_NI_GEOM] == OPTB_G_GROUP ||
            d->node_i[(size_t)c * OPTB_NI_STRIDE + OPTB_NI_GEOM] == OPTB_G_GRID)
          return fail(ctx, -5, "scene: lattice child is not a leaf of the group's subtree");
      }
    }
    if (g == OPTB_G_GROUP || g == OPTB_G_GRID) continue;
    if (ni[OPTB_NI_SKIP] != i + 1) return fail(ctx, -5, "scene: a leaf must skip to the next node");
    if (ni[OPTB_NI_INTER] < OPTB_I_MIRROR || ni[OPTB_NI_INTER] > OPTB_I_PASS) return fail(ctx, -5, "scene: unknown interaction kind");
    if (ni[OPTB_NI_MAT1] < 0 || ni[OPTB_NI_MAT1] >= d->n_materials || ni[OPTB_NI_MAT2] < 0 || ni[OPTB_NI_MAT2] >= d->n_materials)
      return fail(ctx, -5, "scene: material index out of range");
    if (ni[OPTB_NI_CAPSLOT] >= d->n_capslots) return fail(ctx, -5, "scene: cap slot out of range");
    if (ni[OPTB_NI_LEAF] < 0 || ni[OPTB_NI_LEAF] >= d->n_leaves) return fail(ctx, -5, "scene: leaf index out of range");
    auto poly_ok = [&](long long off) {
      if (off < 0 || off + OPTB_POLY_HEADER > d->n_aux) return false;
      const double nv = d->aux[off];
      return nv >= 3 && nv <= 1e6 && off + OPTB_POLY_HEADER + 2 * (long long)nv <= d->n_aux;
    };
    const double* p = d->node_f + (size_t)i * OPTB_NF_STRIDE + OPTB_NF_P;
    if ((g == OPTB_G_POLY2D || g == OPTB_G_POLY3D) && !poly_ok(ni[OPTB_NI_AUX])) return fail(ctx, -5, "scene: polygon record out of range");
    if (g == OPTB_G_ASPHERE && ni[OPTB_NI_AUX] != OPTB_ASPH_PARAMETRIC && ni[OPTB_NI_AUX] != OPTB_ASPH_EXACT_SPH)
      return fail(ctx, -5, "scene: unknown asphere form");
    if (g == OPTB_G_CSG && (int)p[0] == 2) {  // nested composite: a well-formed postfix program in the aux pool
      const long long off = ni[OPTB_NI_AUX];
      if (off < 0 || off + 1 > d->n_aux) return fail(ctx, -5, "scene: CSG program out of range");
      const double nt = d->aux[off];
      if (!(nt >= 3 && nt <= 4096) || off + 1 + 3 * (long long)nt > d->n_aux) return fail(ctx, -5, "scene: CSG program out of range");
      int depth = 0;
      for (int k = 0; k < (int)nt; k++) {
        const int code = (int)d->aux[off + 1 + 3 * k];
        if (code == OPTB_G_CIRCLE || code == OPTB_G_RECT || code == OPTB_G_POLY2D) {
          if (code == OPTB_G_POLY2D && !poly_ok((long long)d->aux[off + 2 + 3 * k])) return fail(ctx, -5, "scene: CSG polygon out of range");
          if (++depth > OPTB_CSG_MAX_DEPTH) return fail(ctx, -5, "scene: CSG program nests too deep");
        } else if (code == OPTB_CSG_SUBTRACT || code == OPTB_CSG_UNION) {
          if (depth < 2) return fail(ctx, -5, "scene: malformed CSG program");
          depth--;
        } else {
          return fail(ctx, -5, "scene: bad CSG token");
        }
      }
      if (depth != 1) return fail(ctx, -5, "scene: malformed CSG program");
    } else if (g == OPTB_G_CSG) {
      if ((int)p[0] != 0 && (int)p[0] != 1) return fail(ctx, -5, "scene: bad CSG op");
      for (int q = 0; q < 2; q++) {
        const int k = (int)p[1 + 3 * q];
        if (k != OPTB_G_CIRCLE && k != OPTB_G_RECT && k != OPTB_G_POLY2D) return fail(ctx, -5, "scene: bad CSG operand");
        if (k == OPTB_G_POLY2D && !poly_ok((long long)p[2 + 3 * q])) return fail(ctx, -5, "scene: CSG polygon out of range");
      }
    }
  }
  cudaSetDevice(ctx->device);
  optb_scene* s = new (std::nothrow) optb_scene();
  if (!s) return -3;
  memset(s, 0, sizeof *s);
  s->n_nodes = d->n_nodes; s->n_leaves = d->n_leaves; s->n_mats = d->n_materials; s->n_mons = d->n_monitors;
  s->n_caps = d->n_capslots; s->n_aux = d->n_aux;
  size_t b_nf = (size_t)d->n_nodes * OPTB_NF_STRIDE * 8, b_ni = (size_t)d->n_nodes * OPTB_NI_STRIDE * 4;
  size_t b_mk = (size_t)d->n_materials * 4, b_mf = (size_t)d->n_materials * OPTB_MF_STRIDE * 8;
  size_t b_mon = (size_t)std::max(d->n_monitors, 1) * kMonBlobStride * 8, b_aux = (size_t)std::max<long long>(d->n_aux, 1) * 8;
  size_t o = 0;
  const size_t b_trav = (size_t)std::max(d->n_nodes, 1) * 64;
  s->off.trav = (uint32_t)o; o = align_up(o + b_trav, 16);
  s->off.nf = (uint32_t)o; o = align_up(o + b_nf, 16);
  s->off.ni = (uint32_t)o; o = align_up(o + b_ni, 16);
  s->off.matk = (uint32_t)o; o = align_up(o + b_mk, 16);
  s->off.matf = (uint32_t)o; o = align_up(o + b_mf, 16);
  s->off.mon = (uint32_t)o; o = align_up(o + b_mon, 16);
  s->off.aux = (uint32_t)o; o = align_up(o + b_aux, 16);
  if (o > 0xfff00000ull) { delete s; return fail(ctx, -6, "scene too large"); }
  s->blob_bytes = (uint32_t)o;
  std::vector<unsigned char> host(o, 0);
  if (d->n_nodes) {
    memcpy(host.data() + s->off.nf, d->node_f, b_nf);
    memcpy(host.data() + s->off.ni, d->node_i, b_ni);
    // parent of every node, from the skip pointers (pre-order: the innermost open group that still covers i)
    std::vector<int32_t> parent(d->n_nodes, -1), open;
    for (int i = 0; i < d->n_nodes; i++) {
      const int32_t* ni = d->node_i + (size_t)i * OPTB_NI_STRIDE;
      while (!open.empty() && d->node_i[(size_t)open.back() * OPTB_NI_STRIDE + OPTB_NI_SKIP] <= i) open.pop_back();
      parent[i] = open.empty() ? -1 : open.back();
      if (ni[OPTB_NI_GEOM] == OPTB_G_GROUP || ni[OPTB_NI_GEOM] == OPTB_G_GRID) open.push_back(i);
    }
    s->cull = new std::vector<unsigned char>();
    std::vector<unsigned char>& cull = *s->cull;
    cull_bits(d, cull);
    for (int i = 0; i < d->n_nodes; i++) {  // compact traversal records
      unsigned char* tv = host.data() + s->off.trav + (size_t)i * 64;
      const int32_t* ni = d->node_i + (size_t)i * OPTB_NI_STRIDE;
      memcpy(tv, d->node_f + (size_t)i * OPTB_NF_STRIDE + OPTB_NF_AABB, 48);
      // box word: bit 0 = test the box (component_group.py:98-107), bit 1 = the box may dismiss by distance (cull_bits)
      const int32_t pack[4] = {ni[OPTB_NI_GEOM], ni[OPTB_NI_SKIP], (ni[OPTB_NI_AABB] ? 1 : 0) | (cull[i] ? 2 : 0), parent[i]};
      memcpy(tv + 48, pack, 16);
    }
  }
  memcpy(host.data() + s->off.matk, d->mat_kind, b_mk);
  memcpy(host.data() + s->off.matf, d->mat_f, b_mf);
  for (int m = 0; m < d->n_monitors; m++) {  // ABI record + the histogram constants derived from it (kMonBlobStride)
    double* mf = (double*)(host.data() + s->off.mon) + (size_t)m * kMonBlobStride;
    memcpy(mf, d->mon_f + (size_t)m * OPTB_MON_STRIDE, OPTB_MON_STRIDE * 8);
    const double w = mf[OPTB_MON_HW] - (-mf[OPTB_MON_HW]), h = mf[OPTB_MON_HH] - (-mf[OPTB_MON_HH]);  // hi - lo as hist_bin forms it
    mf[OPTB_MON_STRIDE] = OPTB_HIST_BINS / w; mf[OPTB_MON_STRIDE + 1] = w / OPTB_HIST_BINS;
    mf[OPTB_MON_STRIDE + 2] = OPTB_HIST_BINS / h; mf[OPTB_MON_STRIDE + 3] = h / OPTB_HIST_BINS;
  }
  for (int i = 0; i < d->n_nodes; i++) asphere_constants((double*)(host.data() + s->off.nf) + (size_t)i * OPTB_NF_STRIDE,
                                                         d->node_i + (size_t)i * OPTB_NI_STRIDE);
  if (d->n_aux) memcpy(host.data() + s->off.aux, d->aux, (size_t)d->n_aux * 8);
  // most children one interaction can emit (decides whether the wavefront machinery is needed)
  int mc = 0;
  for (int i = 0; i < d->n_nodes; i++) {
    const int32_t* ni = d->node_i + (size_t)i * OPTB_NI_STRIDE;
    const double* nf = d->node_f + (size_t)i * OPTB_NF_STRIDE;
    if (ni[OPTB_NI_AABB]) s->has_boxes = 1;
    if (ni[OPTB_NI_GEOM] == OPTB_G_ASPHERE) s->has_asph = 1;
    if (ni[OPTB_NI_GEOM] == OPTB_G_GRID) s->has_grid = 1;
    if (ni[OPTB_NI_INTER] == OPTB_I_PASS) s->has_pass = 1;
    int k = 0;
    switch (ni[OPTB_NI_INTER]) {
      case OPTB_I_MIRROR: k = (nf[OPTB_NF_REFL] > 0) + (nf[OPTB_NF_TRANS] > 0); break;
      case OPTB_I_REFRACT: k = nf[OPTB_NF_REFL] > 0 ? 2 : 1; break;
      case OPTB_I_THINLENS: case OPTB_I_PASS: k = 1; break;
      default: k = 0;
    }
    mc = std::max(mc, k);
  }
  s->max_children = mc;
  size_t hist_bytes = (size_t)d->n_monitors * OPTB_HIST_BINS * (OPTB_HIST_BINS + 1) * 4;
  size_t budget = ctx->smem_optin > 2048 ? ctx->smem_optin - 2048 : 0;
  s->in_smem = (o <= budget);
  size_t used = s->in_smem ? o : 0;
  s->hist_smem = (hist_bytes > 0 && used + hist_bytes <= std::min<size_t>(budget, used + 65536));
  s->smem_bytes = (uint32_t)(used + (s->hist_smem ? hist_bytes : 0));
  cudaError_t e = cudaSuccess;
  int pick = -1;  // smallest retired blob that is large enough
  for (int k = 0; k < 8; k++)
    if (ctx->pool[k].p && ctx->pool[k].cap >= o && (pick < 0 || ctx->pool[k].cap < ctx->pool[pick].cap)) pick = k;
  if (pick >= 0) {
    s->d_blob = ctx->pool[pick].p; s->blob_cap = ctx->pool[pick].cap;
    ctx->pool[pick].p = nullptr; ctx->pool[pick].cap = 0;
  } else {
    s->blob_cap = (o + 65535) & ~(size_t)65535;
    e = cudaMalloc((void**)&s->d_blob, s->blob_cap);
    if (e != cudaSuccess) { delete s; return fail(ctx, -10, "cudaMalloc(scene)", e); }
  }
  e = cudaMemcpy(s->d_blob, host.data(), o, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->last_use, cudaEventDisableTiming);
  if (e != cudaSuccess) { cudaFree(s->d_blob); delete s; return fail(ctx, -10, "cudaMemcpy(scene)", e); }
  *out = s;
  return 0;
}

extern "C" int optb_scene_update_nodes(optb_ctx* ctx, optb_scene* s, const optb_scene_desc* d, const int32_t* nodes,
                                       int32_t n, void* stream_v) {
  if (!ctx || !s || !d || (n > 0 && !nodes)) return -1;
  if (d->n_nodes != s->n_nodes || d->n_leaves != s->n_leaves || d->n_materials != s->n_mats || d->n_aux != s->n_aux ||
      d->n_capslots != s->n_caps)
    return fail(ctx, -5, "optb_scene_update_nodes: the tables no longer have the shape of the uploaded scene");
  cudaSetDevice(ctx->device);
  cudaStream_t st = (cudaStream_t)stream_v;
  for (int k = 0; k < n; k++) {
    const int i = nodes[k];
    if (i < 0 || i >= d->n_nodes) return fail(ctx, -5, "optb_scene_update_nodes: node index out of range");
    const int32_t* ni = d->node_i + (size_t)i * OPTB_NI_STRIDE;
    const double* nf = d->node_f + (size_t)i * OPTB_NF_STRIDE;
    double row[OPTB_NF_STRIDE];
    memcpy(row, nf, sizeof row);
    asphere_constants(row, ni);
    // (pageable source: the call returns once `row` has been copied to the driver's staging memory)
    CK(cudaMemcpyAsync(s->d_blob + s->off.nf + (size_t)i * OPTB_NF_STRIDE * 8, row, OPTB_NF_STRIDE * 8, cudaMemcpyHostToDevice, st), "update node_f");
    CK(cudaMemcpyAsync(s->d_blob + s->off.ni + (size_t)i * OPTB_NI_STRIDE * 4, ni, OPTB_NI_STRIDE * 4, cudaMemcpyHostToDevice, st), "update node_i");
    CK(cudaMemcpyAsync(s->d_blob + s->off.trav + (size_t)i * 64, nf + OPTB_NF_AABB, 48, cudaMemcpyHostToDevice, st), "update box");
  }
  // a moved node may change which boxes are allowed to dismiss by distance, up to the root: recompute, send the changes
  if (s->cull) {
    std::vector<unsigned char> now;
    cull_bits(d, now);
    for (int i = 0; i < d->n_nodes; i++) {
      if (now[i] == (*s->cull)[i]) continue;
      const int32_t word = (d->node_i[(size_t)i * OPTB_NI_STRIDE + OPTB_NI_AABB] ? 1 : 0) | (now[i] ? 2 : 0);
      CK(cudaMemcpyAsync(s->d_blob + s->off.trav + (size_t)i * 64 + 56, &word, 4, cudaMemcpyHostToDevice, st), "update box word");
    }
    s->cull->swap(now);
  }
  // scene-wide properties that pick the kernel variant follow the new rows
  int mc = 0, boxes = 0, asph = 0, pass = 0;
  for (int i = 0; i < d->n_nodes; i++) {
    const int32_t* ni = d->node_i + (size_t)i * OPTB_NI_STRIDE;
    const double* nf = d->node_f + (size_t)i * OPTB_NF_STRIDE;
    if (ni[OPTB_NI_AABB]) boxes = 1;
    if (ni[OPTB_NI_GEOM] == OPTB_G_ASPHERE) asph = 1;
    if (ni[OPTB_NI_INTER] == OPTB_I_PASS) pass = 1;
    int k = 0;
    switch (ni[OPTB_NI_INTER]) {
      case OPTB_I_MIRROR: k = (nf[OPTB_NF_REFL] > 0) + (nf[OPTB_NF_TRANS] > 0); break;
      case OPTB_I_REFRACT: k = nf[OPTB_NF_REFL] > 0 ? 2 : 1; break;
      case OPTB_I_THINLENS: case OPTB_I_PASS: k = 1; break;
      default: k = 0;
    }
    mc = std::max(mc, k);
  }
  s->max_children = mc; s->has_boxes = boxes; s->has_asph = asph; s->has_pass = pass;
  return 0;
}

extern "C" int optb_scene_destroy(optb_ctx* ctx, optb_scene* s) {
  if (!s) return 0;
  if (ctx) cudaSetDevice(ctx->device);
  if (s->d_blob) {
    int slot = -1;
    if (ctx && s->blob_cap <= ((size_t)64 << 20))
      for (int k = 0; k < 8 && slot < 0; k++) if (!ctx->pool[k].p) slot = k;
    if (s->used) cudaEventSynchronize(s->last_use);  // no kernel may still be reading the blob
    if (slot >= 0) {
      ctx->pool[slot].p = s->d_blob; ctx->pool[slot].cap = s->blob_cap;
    } else {
      cudaFree(s->d_blob);
    }
  }
  if (s->last_use) cudaEventDestroy(s->last_use);
  delete s->cull;
  delete s;
  return 0;
}

namespace {
struct WsLayout {
  size_t hdr, w, c, nchild, gen_first, gen_last, sums, iota, perm, key_sorted, key_dense, slot_a, slot_b, cub, cub_bytes, total;
  long long cap;  // wavefront capacity (rays)
};
size_t raybuf_bytes(long long cap) { return align_up((size_t)cap * sizeof(RayRec), 256) + 2 * align_up((size_t)cap * 8, 256); }
WsLayout ws_layout(long long n_rays, long long max_live, bool split) {
  WsLayout L{};
  size_t o = 0;
  L.hdr = o; o += align_up(sizeof(Header), 256);
  L.cap = split ? std::max(max_live, n_rays) : 0;
  if (split) {
    L.w = o; o += raybuf_bytes(2 * L.cap);  // two sparse child buffers, used alternately as source and destination
    L.c = o; o += raybuf_bytes(2 * L.cap);
    L.slot_a = o; o += align_up((size_t)L.cap * 4, 256);
    L.slot_b = o; o += align_up((size_t)L.cap * 4, 256);
    L.key_dense = o; o += align_up((size_t)L.cap * 4, 256);
    L.nchild = o; o += align_up((size_t)L.cap, 256);
    L.gen_first = o; o += align_up((size_t)n_rays * 4, 256);
    L.gen_last = o; o += align_up((size_t)n_rays * 4, 256);
    L.sums = o; o += align_up(((size_t)L.cap / kTile + 2) * 4, 256);
    L.iota = o; o += align_up((size_t)L.cap * 4, 256);
    L.perm = o; o += align_up((size_t)L.cap * 4, 256);
    L.key_sorted = o; o += align_up((size_t)L.cap * 4, 256);
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (int)std::min<long long>(L.cap, 0x7fffffffll), 0, 32);
    L.cub = o; L.cub_bytes = align_up(tb + 256, 256); o += L.cub_bytes;
  }
  L.total = o;
  return L;
}
RayBuf make_raybuf(unsigned char* base, long long cap) {
  RayBuf b;
  b.rec = (RayRec*)base;
  b.key = (uint2*)(base + align_up((size_t)cap * sizeof(RayRec), 256));
  b.rk = (uint2*)(base + align_up((size_t)cap * sizeof(RayRec), 256) + align_up((size_t)cap * 8, 256));
  return b;
}
bool needs_wavefront(const optb_scene* s, const optb_params* p) {
  return s->max_children > 1 || p->chain_len > 0 || p->flag_ambiguity || p->reference_roots || s->has_pass;  // (these run the SPLIT variants)
}
// Interact caps make the result depend on the reference's sequential order as soon as two rays of one family can
// be in flight: splitting scenes, or several initial rays sharing an `_id` (family column given).
bool needs_serial(const optb_scene* s, const optb_rays* r, const optb_params* p) {
  return s->n_caps > 0 && !p->caps_slack && (s->max_children > 1 || r->family != nullptr);
}
struct SerialLayout { size_t hdr, off, cursor, roots, ring, total; };
SerialLayout serial_layout(long long n_rays, long long n_fam, long long ring_entries) {
  SerialLayout L{};
  size_t o = 0;
  L.hdr = o; o += align_up(sizeof(Header), 256);
  L.off = o; o += align_up((size_t)(n_fam + 2) * 4, 256);
  L.cursor = o; o += align_up((size_t)(n_fam + 1) * 4, 256);
  L.roots = o; o += align_up((size_t)std::max<long long>(n_rays, 1) * 4, 256);
  L.ring = o; o += raybuf_bytes(std::max<long long>(ring_entries, 1));
  L.total = o;
  return L;
}
}  // namespace

extern "C" int64_t optb_workspace_bytes(const optb_scene* scene, int64_t n_rays, int64_t max_live) {
  if (!scene || n_rays < 0) return -1;
  // the caller may later pass chain_len > 0, so always size for the wavefront when asked for max_live > 0
  bool split = scene->max_children > 1 || max_live > 0 || scene->has_pass;
  int64_t need = (int64_t)ws_layout(n_rays, std::max<int64_t>(max_live, n_rays), split).total;
  if (scene->n_caps > 0)  // family-serial mode: max_live = total FIFO entries over all families
    need = std::max<int64_t>(need, (int64_t)serial_layout(n_rays, n_rays, std::max<int64_t>(max_live, 8 * n_rays)).total);
  return need;
}

static int trace_impl(optb_ctx* ctx, const optb_scene* scene, const optb_rays* rays, const optb_params* prm,
                      optb_result* out, void* workspace, int64_t workspace_bytes, cudaStream_t st, uint32_t root_base,
                      bool zero_hist);

extern "C" int optb_trace(optb_ctx* ctx, const optb_scene* scene, const optb_rays* rays, const optb_params* prm,
                          optb_result* out, void* workspace, int64_t workspace_bytes, void* stream_v) {
  if (!ctx || !scene || !rays || !prm || !out) return -1;
  return trace_impl(ctx, scene, rays, prm, out, workspace, workspace_bytes, (cudaStream_t)stream_v, 0u, true);
}

static int trace_impl(optb_ctx* ctx, const optb_scene* scene, const optb_rays* rays, const optb_params* prm,
                      optb_result* out, void* workspace, int64_t workspace_bytes, cudaStream_t st, uint32_t root_base,
                      bool zero_hist) {
  cudaSetDevice(ctx->device);
  if (!out->counters) return fail(ctx, -7, "result.counters is required");
  if (rays->n >= 0xffffffffll) return fail(ctx, -7, "at most 2^32-1 rays per call");
  if (prm->record_hits && out->hit_capacity > 0 && !out->hit_monitor && !out->hit_key) return fail(ctx, -7, "record_hits needs hit_monitor or hit_key");
  if (prm->record_hits && out->hit_key && (prm->max_trace_num > (1ll << 24) || scene->n_mons > 256))
    return fail(ctx, -7, "hit_key packs pop into 24 bits and the monitor into 8: max_trace_num <= 2^24 and <= 256 monitors");
  if (prm->flag_ambiguity && !out->root_flags) return fail(ctx, -7, "flag_ambiguity needs result.root_flags [rays.n]");
  if (scene->n_caps > 0 && (!out->cap_counts || prm->n_families < 1)) return fail(ctx, -7, "scene has interact caps: cap_counts/n_families required");
  // cap_counts is [n_capslots][n_families]; without a family column every initial ray is its own family (column
  // index = ray index), so the table must have a column per ray
  if (scene->n_caps > 0 && !rays->family && (long long)prm->n_families < rays->n)
    return fail(ctx, -7, "scene has interact caps and rays.family is NULL: n_families must be >= rays.n");
  const bool serial = needs_serial(scene, rays, prm);
  if (serial && prm->flag_ambiguity) return fail(ctx, -7, "flag_ambiguity is not available on the family-serial path (binding interact caps)");
  const bool split = !serial && needs_wavefront(scene, prm);
  WsLayout L = ws_layout(rays->n, 0, false);
  long long cap = 0;
  SerialLayout SL{};
  long long n_fam = 0, qcap = 0;
  if (serial) {
    n_fam = rays->family ? prm->n_families : rays->n;
    if (n_fam < 1) return fail(ctx, -7, "n_families must be >= 1");
    const long long want = std::min<long long>(prm->max_trace_num + 2, 1ll << 20);
    const size_t fixed = serial_layout(rays->n, n_fam, 1).total;
    if ((size_t)workspace_bytes <= fixed) return fail(ctx, -8, "workspace too small (see optb_workspace_bytes)");
    long long lo = 1, hi = std::max<long long>(want, 2) * n_fam;
    while (lo < hi) {
      long long mid = lo + (hi - lo + 1) / 2;
      if ((int64_t)serial_layout(rays->n, n_fam, mid).total <= workspace_bytes) lo = mid; else hi = mid - 1;
    }
    qcap = std::min<long long>(lo / n_fam, want);
    if (qcap < 4) return fail(ctx, -8, "workspace too small for the per-family FIFO (see optb_workspace_bytes)");
    SL = serial_layout(rays->n, n_fam, qcap * n_fam);
    L.hdr = SL.hdr;
  } else if (split) {
    // find the largest wavefront capacity the given workspace supports
    if (workspace_bytes < (int64_t)ws_layout(rays->n, rays->n, true).total) return fail(ctx, -8, "workspace too small (see optb_workspace_bytes)");
    // upper end of the search: what the workspace could hold at ~100 B per live ray (far above the real cost of
    // ~540 B), so the capacity is set by the caller's workspace and never by the size of the batch: one initial ray
    // through cascaded beam splitters may need thousands of live rays (the reference allows 2000 pops)
    long long lo = rays->n, hi = std::max<long long>(rays->n, (long long)(workspace_bytes / 100));
    hi = std::min<long long>(hi, 0x7ffffff0ll);
    while (lo < hi) {
      long long mid = lo + (hi - lo + 1) / 2;
      if ((int64_t)ws_layout(rays->n, mid, true).total <= workspace_bytes) lo = mid; else hi = mid - 1;
    }
    cap = lo;
    L = ws_layout(rays->n, cap, true);
  } else if (workspace_bytes < (int64_t)L.total) {
    return fail(ctx, -8, "workspace too small (see optb_workspace_bytes)");
  }
  unsigned char* ws = (unsigned char*)workspace;
  Header* hdr = (Header*)(ws + L.hdr);

  CK(cudaMemsetAsync(out->counters, 0, sizeof(int64_t) * OPTB_C_COUNT, st), "memset counters");
  CK(cudaMemsetAsync(hdr, 0, sizeof(Header), st), "memset header");
  if (prm->record_hist && scene->n_mons && !out->hist_y) return fail(ctx, -7, "record_hist needs hist_y and hist_yz");
  if (prm->record_hist && scene->n_mons && zero_hist) {
    if (!out->hist_y || !out->hist_yz) return fail(ctx, -7, "record_hist needs hist_y and hist_yz");
    CK(cudaMemsetAsync(out->hist_y, 0, sizeof(int64_t) * OPTB_HIST_BINS * scene->n_mons, st), "memset hist");
    CK(cudaMemsetAsync(out->hist_yz, 0, sizeof(int64_t) * OPTB_HIST_BINS * OPTB_HIST_BINS * scene->n_mons, st), "memset hist");
  }

  if (prm->flag_ambiguity) CK(cudaMemsetAsync(out->root_flags, 0, sizeof(uint32_t) * (size_t)rays->n, st), "memset root flags");
  TraceArgs a;
  memset(&a, 0, sizeof a);
  a.blob = scene->d_blob; a.blob_bytes = scene->blob_bytes; a.off = scene->off;
  a.n_nodes = scene->n_nodes; a.n_mons = scene->n_mons;
  a.in0 = *rays; a.gen0 = 1; a.n_in = rays->n; a.n_in_dev = nullptr;
  a.max_trace = prm->max_trace_num; a.unit = prm->unit;
  a.rec_seg = prm->record_segments; a.rec_hit = prm->record_hits; a.rec_hist = prm->record_hist && scene->n_mons > 0;
  a.chain_len = prm->chain_len; a.n_families = prm->n_families;
  a.fam_shared = (rays->family != nullptr) || prm->caps_slack;
  a.hist_smem = a.rec_hist && scene->hist_smem;
  a.out = *out;
  a.counters = (unsigned long long*)out->counters;
  a.hdr = hdr;
  a.root_base = root_base;
  a.has_boxes = scene->has_boxes;
  if (split) {
    a.w = make_raybuf(ws + L.w, 2 * cap);
    a.c = make_raybuf(ws + L.c, 2 * cap);
    a.nchild = ws + L.nchild;
    a.gen_first = (uint32_t*)(ws + L.gen_first);
    a.gen_last = (uint32_t*)(ws + L.gen_last);
  }
  uint32_t smem = scene->in_smem ? scene->smem_bytes : (a.hist_smem ? scene->smem_bytes : 0);
  using Kern = void (*)(const TraceArgs);
  // box mode: 0 none, 1 pre-order walk with culling, 2 walk + lattice windows (scenes with OPTB_G_GRID groups)
  const int boxmode = !scene->has_boxes ? 0 : ((scene->has_grid && OPTB_GRID_WALK) ? 2 : 1);
  // [smem][box mode][aspheres][split] for the parallel path; the family-serial path keeps one general variant per
  // staging mode
#define OPTB_K(S, B, A, P) trace_kernel<S, false, B, A, P>
#define OPTB_KROW(S, B) {{OPTB_K(S, B, false, false), OPTB_K(S, B, false, true)}, {OPTB_K(S, B, true, false), OPTB_K(S, B, true, true)}}
  static const Kern table[2][3][2][2] = {{OPTB_KROW(false, 0), OPTB_KROW(false, 1), OPTB_KROW(false, 2)},
                                         {OPTB_KROW(true, 0), OPTB_KROW(true, 1), OPTB_KROW(true, 2)}};
#undef OPTB_KROW
#undef OPTB_K
  static const Kern serial_table[2] = {trace_kernel<false, true, 1, true, true>, trace_kernel<true, true, 1, true, true>};
  // diagnostics (params.flag_ambiguity): the general variant of the parallel path + the A.9 pass at every pop
  static const Kern flag_table[2] = {trace_kernel<false, false, 1, true, true, true>, trace_kernel<true, false, 1, true, true, true>};
  Kern kern = serial ? serial_table[scene->in_smem ? 1 : 0]
                     : table[scene->in_smem ? 1 : 0][boxmode][scene->has_asph ? 1 : 0][split ? 1 : 0];
  // a pass-through leaf (OPTB_I_PASS) only exists in the general variants: this one, or the three kinds below
  static const Kern pass_table[2] = {trace_kernel<false, false, 1, true, true, false, false, true>,
                                     trace_kernel<true, false, 1, true, true, false, false, true>};
  if (scene->has_pass && !serial) kern = pass_table[scene->in_smem ? 1 : 0];
  if (prm->flag_ambiguity) kern = flag_table[scene->in_smem ? 1 : 0];
  // params.reference_roots: the general variants with brentq_dev (parallel and family-serial)
  static const Kern brent_table[2][2] = {{trace_kernel<false, false, 1, true, true, false, true>, trace_kernel<true, false, 1, true, true, false, true>},
                                         {trace_kernel<false, true, 1, true, true, false, true>, trace_kernel<true, true, 1, true, true, false, true>}};
  if (prm->reference_roots) {
    if (prm->flag_ambiguity) return fail(ctx, -7, "reference_roots and flag_ambiguity are separate diagnostics modes");
    kern = brent_table[serial ? 1 : 0][scene->in_smem ? 1 : 0];
  }
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<uint32_t>(smem, 1024)), "smem attr");
  int occ = 1;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kBlock, smem), "occupancy");
  if (occ < 1) occ = 1;
  const int full_grid = ctx->sm_count * occ;

  unsigned long long gens = 0, launches = 0;
  long long n_in = rays->n;
  if (serial && n_in > 0) {
    unsigned int* off = (unsigned int*)(ws + SL.off);
    unsigned int* cursor = (unsigned int*)(ws + SL.cursor);
    uint32_t* roots = (uint32_t*)(ws + SL.roots);
    CK(cudaMemsetAsync(off, 0, (size_t)(n_fam + 2) * 4, st), "memset family offsets");
    CK(cudaMemsetAsync(cursor, 0, (size_t)(n_fam + 1) * 4, st), "memset family cursors");
    const int g = (int)std::min<long long>(full_grid * 4, (n_in + 255) / 256);
    fam_count_kernel<<<g, 256, 0, st>>>(rays->family, n_in, off);
    scan_sums_kernel<<<1, 1024, 0, st>>>(off, (int)(n_fam + 1), hdr, a.counters, 0xffffffffu);  // exclusive, in place
    fam_scatter_kernel<<<g, 256, 0, st>>>(rays->family, n_in, off, cursor, roots);
    a.fam_off = off; a.fam_roots = roots; a.qcap = (uint32_t)qcap;
    a.w = make_raybuf(ws + SL.ring, qcap * n_fam);
    a.fam_shared = 0;
    a.n_in = n_fam;
    void* kargs[] = {(void*)&a};
    int grid = (int)std::min<long long>(full_grid, (n_fam + kBlock - 1) / kBlock);
    CK(cudaLaunchKernel((const void*)kern, dim3(grid), dim3(kBlock), kargs, smem, st), "launch trace_kernel (family-serial)");
    launches += 4; gens = 1;
    n_in = 0;
  }
  int key_bits = 1;
  while ((1ll << key_bits) < 2ll * std::max(scene->n_leaves, 1)) key_bits++;
  while (n_in > 0) {
    if (split && !a.gen0 && n_in >= 2048 && n_in < 0x7fffffffll) {
      uint32_t* iota = (uint32_t*)(ws + L.iota);
      uint32_t* perm = (uint32_t*)(ws + L.perm);
      uint32_t* key_dense = (uint32_t*)(ws + L.key_dense);
      sort_prep_kernel<<<std::min(full_grid * 4, (int)((n_in + 255) / 256)), 256, 0, st>>>(iota, key_dense, a.w.key, a.w.rk, a.slot, n_in, key_bits);
      size_t tb = L.cub_bytes;
      CK(cub::DeviceRadixSort::SortPairs(ws + L.cub, tb, (const uint32_t*)key_dense, (uint32_t*)(ws + L.key_sorted),
                                         (const uint32_t*)iota, perm, (int)n_in, 0, key_bits + 1, st), "radix sort");
      a.perm = perm;
      launches += 4;
    } else {
      a.perm = nullptr;
    }
    // Small wavefronts: the host round trip per generation (read the header, synchronise) costs more than the
    // generation itself, so up to kUnseen further generations are enqueued without knowing their size. They take
    // it from the header (at most twice the previous one, which bounds the grids) and are no-ops once the
    // wavefront has died. Sorting needs the size on the host; the bound keeps these below the sort threshold.
    constexpr int kUnseen = 8;
    int unseen_left = 0;
    if (split) {
      long long bound = n_in;
      while (unseen_left < kUnseen && 2 * bound < 2048 && 2 * bound <= cap) { bound *= 2; unseen_left++; }
    }
    long long bound = n_in;
    for (int u = 0; u <= unseen_left; u++) {
      const bool unseen = u > 0;
      if (unseen) { bound *= 2; a.perm = nullptr; }
      long long want = (bound + kBlock - 1) / kBlock;
      int grid = (int)std::min<long long>(full_grid, want);
      a.n_in = bound;
      a.n_in_dev = unseen ? &hdr->n_next : nullptr;
      void* kargs[] = {(void*)&a};
      CK(cudaLaunchKernel((const void*)kern, dim3(grid), dim3(kBlock), kargs, smem, st), "launch trace_kernel");
      launches++; gens++;
      if (!split) break;
      int ntiles = (int)((bound + kTile - 1) / kTile);
      unsigned int* sums = (unsigned int*)(ws + L.sums);
      tile_sums_kernel<<<ntiles, kScanBlock, 0, st>>>(a.nchild, bound, sums, unseen ? hdr : nullptr);
      scan_sums_kernel<<<1, 1024, 0, st>>>(sums, ntiles, hdr, a.counters, (unsigned int)std::min<long long>(cap, 0xffffffffll), unseen);
      // the children just written to a.c become the next generation's source; the other buffer is free again
      uint32_t* slot_next = (uint32_t*)(ws + ((gens & 1) ? L.slot_a : L.slot_b));
      slots_kernel<<<ntiles, kScanBlock, 0, st>>>(a.nchild, bound, sums, slot_next, hdr, unseen);
      const int mgrid = std::min(full_grid * 2, std::max(1, (int)((2 * bound + 255) / 256)));
      mark_kernel<<<mgrid, 256, 0, st>>>(a.c.key, slot_next, hdr, (uint32_t*)a.gen_first, (uint32_t*)a.gen_last);
      rank_kernel<<<mgrid, 256, 0, st>>>(a.c.key, a.c.rk, slot_next, hdr, a.gen_first, a.gen_last);
      std::swap(a.w, a.c);
      a.slot = slot_next;
      launches += 5;
      a.gen0 = 0;
    }
    if (!split) break;
    a.n_in_dev = nullptr;
    CK(cudaMemcpyAsync(ctx->h_hdr, hdr, sizeof(Header), cudaMemcpyDeviceToHost, st), "read header");
    CK(cudaStreamSynchronize(st), "sync generation");
    n_in = ((Header*)ctx->h_hdr)->n_next;
  }
  if (prm->flag_ambiguity && rays->n > 0) {
    count_flags_kernel<<<(int)std::min<long long>((rays->n + 255) / 256, (long long)full_grid * 4), 256, 0, st>>>(out->root_flags, rays->n, a.counters);
    launches++;
  }
  finish_kernel<<<1, 1, 0, st>>>(a.counters, gens, launches + 1, hdr);
  CK(cudaGetLastError(), "kernel launch");
  cudaStreamCaptureStatus capturing = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(st, &capturing);
  if (capturing == cudaStreamCaptureStatusNone) {  // (a captured event could not be waited on when the scene is released;
    CK(cudaEventRecord(scene->last_use, st), "record scene use");  //  the owner of a graph synchronises before that)
    const_cast<optb_scene*>(scene)->used = true;
  }
  return 0;
}

namespace {
struct ArenaCursor {
  unsigned char* base; size_t off, cap;
  void* take(size_t bytes) { size_t o = align_up(off, 256); off = o + bytes; return off <= cap ? base + o : nullptr; }
};
}  // namespace

static int trace_host_once(optb_ctx* ctx, const optb_scene* scene, const optb_rays* rays, const optb_params* prm,
                           optb_result* out, int64_t live);

// Chunked, three-stream version of optb_trace_host for scenes that cannot split: the host->device copy of chunk
// c+1, the trace of chunk c and the device->host copy of chunk c-1 overlap (PCIe is full duplex). Rows of chunk c
// land behind the rows of chunks < c, so the host result is the same set of rows as the one-shot path; row keys
// stay global through root_base. Returns 1 when a chunk overflowed its (estimated) device capacity: the caller
// then repeats the whole batch on the one-shot path.
static int trace_host_pipelined(optb_ctx* ctx, const optb_scene* scene, const optb_rays* rays, const optb_params* prm,
                                optb_result* out) {
  cudaSetDevice(ctx->device);
#ifndef OPTB_HOST_CHUNK
#define OPTB_HOST_CHUNK (1 << 20)
#endif
  constexpr int64_t kChunk = OPTB_HOST_CHUNK;
  const int64_t n = rays->n;
  const int nch = (int)((n + kChunk - 1) / kChunk);
  const int64_t segcap = prm->record_segments ? out->seg_capacity : 0, hitcap = prm->record_hits ? out->hit_capacity : 0;
  auto chunk_cap = [&](int64_t total) { return total ? std::min<int64_t>(total, (int64_t)((double)total * kChunk / n * 1.25) + 4096) : 0; };
  const int64_t cseg = chunk_cap(segcap), chit = chunk_cap(hitcap);
  const int64_t wsb = (int64_t)ws_layout(kChunk, 0, false).total;
  if (!ctx->s_h2d) {
    CK(cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking), "stream");
    CK(cudaStreamCreateWithFlags(&ctx->s_run, cudaStreamNonBlocking), "stream");
    CK(cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking), "stream");
  }
  if (ctx->h_chunk_n < (size_t)nch) {
    if (ctx->h_chunk) cudaFreeHost(ctx->h_chunk);
    ctx->h_chunk = nullptr; ctx->h_chunk_n = 0;
    CK(cudaMallocHost((void**)&ctx->h_chunk, sizeof(unsigned long long) * OPTB_C_COUNT * nch), "pinned counters");
    ctx->h_chunk_n = nch;
  }
  const int64_t sortb = prm->sorted_rows ? optb_sort_workspace_bytes(std::max(cseg, chit)) : 0;
  size_t need = 256 * 128 + (size_t)n * (8 * 13 + 8) + (prm->sorted_rows ? 4 : 2) * ((size_t)cseg * (13 * 8 + 16) + (size_t)chit * (10 * 8 + 12 + 8) + 40 * 256) +
                (size_t)std::max(scene->n_mons, 1) * OPTB_HIST_BINS * (OPTB_HIST_BINS + 1) * 8 + 2 * (64 * 8 + (size_t)wsb) +
                (size_t)sortb + 4096;
  if (ctx->arena_bytes < need) {
    if (ctx->arena) cudaFree(ctx->arena);
    ctx->arena = nullptr; ctx->arena_bytes = 0;
    CK(cudaMalloc(&ctx->arena, need), "cudaMalloc(arena)");
    ctx->arena_bytes = need;
  }
  ArenaCursor ac{(unsigned char*)ctx->arena, 0, ctx->arena_bytes};
  // full-length device input columns; chunks are slices of them
  optb_rays dr = *rays;
  const double* const* src_f = &rays->ox;
  const double** dst_f = &dr.ox;
  for (int f = 0; f < kRayF64; f++) {
    if (!src_f[f]) { dst_f[f] = nullptr; continue; }
    const size_t rows = ((rays->broadcast >> f) & 1u) ? 1 : (size_t)n;
    dst_f[f] = (const double*)ac.take(rows * 8);
    if (!dst_f[f]) return fail(ctx, -9, "arena sizing");
    if (rows == 1) CK(cudaMemcpyAsync((void*)dst_f[f], src_f[f], 8, cudaMemcpyHostToDevice, ctx->s_h2d), "H2D rays");
  }
  if (rays->flags) dr.flags = (const uint32_t*)ac.take((size_t)n * 4);
  if (rays->family) dr.family = (const int32_t*)ac.take((size_t)n * 4);
  // dev: the two result sets the traces of even / odd chunks append to; srt (sorted_rows): the same rows in reference
  // order, gathered there right behind the trace, so that the copy back never waits for a later chunk's trace
  struct Col { void* dev[2]; void* srt[2]; unsigned char* host; size_t elt; int kind; };
  std::vector<Col> cols;
  optb_result dv[2] = {*out, *out}, sv[2] = {*out, *out};
  const bool sorted = prm->sorted_rows != 0;
  auto add = [&](size_t field_off, void* host_ptr, size_t elt, int64_t capn, int kind) {
    for (int s = 0; s < 2; s++) *(void**)((unsigned char*)&dv[s] + field_off) = *(void**)((unsigned char*)&sv[s] + field_off) = nullptr;
    if (!host_ptr || capn == 0) return;
    Col c{{ac.take((size_t)capn * elt), ac.take((size_t)capn * elt)}, {nullptr, nullptr}, (unsigned char*)host_ptr, elt, kind};
    if (sorted) { c.srt[0] = ac.take((size_t)capn * elt); c.srt[1] = ac.take((size_t)capn * elt); }
    for (int s = 0; s < 2; s++) {
      *(void**)((unsigned char*)&dv[s] + field_off) = c.dev[s];
      *(void**)((unsigned char*)&sv[s] + field_off) = c.srt[s];
    }
    cols.push_back(c);
  };
#define OPTB_FIELD(name) offsetof(optb_result, name), (void*)out->name
  {
    const size_t seg0 = offsetof(optb_result, seg_ox), hit0 = offsetof(optb_result, hit_px);
    double* const* seg_h = &out->seg_ox; double* const* hit_h = &out->hit_px;
    for (int f = 0; f < 13; f++) add(seg0 + 8 * f, seg_h[f], 8, cseg, 0);
    add(OPTB_FIELD(seg_flags), 4, cseg, 0); add(OPTB_FIELD(seg_root), 4, cseg, 0);
    add(OPTB_FIELD(seg_pop), 4, cseg, 0); add(OPTB_FIELD(seg_leaf), 4, cseg, 0);
    add(OPTB_FIELD(hit_monitor), 4, chit, 1); add(OPTB_FIELD(hit_root), 4, chit, 1); add(OPTB_FIELD(hit_pop), 4, chit, 1);
    for (int f = 0; f < 10; f++) add(hit0 + 8 * f, hit_h[f], 8, chit, 1);
    add(OPTB_FIELD(hit_key), 8, chit, 1);
  }
#undef OPTB_FIELD
  for (auto& c : cols) if (!c.dev[0] || !c.dev[1] || (sorted && (!c.srt[0] || !c.srt[1]))) return fail(ctx, -9, "arena sizing");
  int end_bit = 33;  // keys are root << 32 | ...: sort only the bits a root index of this batch can occupy
  while (end_bit < 64 && (1ull << (end_bit - 32)) <= (unsigned long long)n) end_bit++;
  const size_t hy = (size_t)std::max(scene->n_mons, 1) * OPTB_HIST_BINS * 8, hyz = hy * OPTB_HIST_BINS;
  int64_t* d_hy = (int64_t*)ac.take(hy); int64_t* d_hyz = (int64_t*)ac.take(hyz);
  void* ws[2];
  for (int s = 0; s < 2; s++) {
    dv[s].seg_capacity = cseg; dv[s].hit_capacity = chit;
    dv[s].hist_y = d_hy; dv[s].hist_yz = d_hyz; dv[s].cap_counts = nullptr;
    dv[s].counters = (int64_t*)ac.take(OPTB_C_COUNT * 8);
    ws[s] = ac.take((size_t)wsb);
    if (!ws[s] || !dv[s].counters) return fail(ctx, -9, "arena sizing");
    dv[s].root_flags = nullptr;
  }
  void* sort_ws = sortb ? ac.take((size_t)sortb) : nullptr;
  if (sortb && !sort_ws) return fail(ctx, -9, "arena sizing");
  if (prm->record_hist && scene->n_mons) {
    CK(cudaMemsetAsync(d_hy, 0, hy, ctx->s_run), "memset hist");
    CK(cudaMemsetAsync(d_hyz, 0, hyz, ctx->s_run), "memset hist");
  }
  struct Events {  // destroyed on every exit path
    std::vector<cudaEvent_t> v;
    explicit Events(int n) : v(n) { for (auto& e : v) cudaEventCreateWithFlags(&e, cudaEventDisableTiming); }
    ~Events() { for (auto& e : v) cudaEventDestroy(e); }
    cudaEvent_t& operator[](int i) { return v[i]; }
  } ev_h2d(nch), ev_run(nch), ev_d2h(nch);
  unsigned long long tot[OPTB_C_COUNT] = {0};
  int64_t seg_off = 0, hit_off = 0;
  bool chunk_overflow = false;
  int rc = 0;
  auto drain = [&](int k) -> int {
    CK(cudaEventSynchronize(ev_run[k]), "sync chunk");
    const unsigned long long* hc = ctx->h_chunk + (size_t)k * OPTB_C_COUNT;
    if (hc[OPTB_C_STATUS] & (OPTB_ST_SEG_OVERFLOW | OPTB_ST_HIT_OVERFLOW)) chunk_overflow = true;
    int64_t rows[2] = {std::min<int64_t>((int64_t)hc[OPTB_C_SEGMENTS], cseg), std::min<int64_t>((int64_t)hc[OPTB_C_HITS], chit)};
    int64_t room[2] = {segcap - seg_off, hitcap - hit_off};
    for (int kind = 0; kind < 2; kind++)
      if (rows[kind] > room[kind]) { rows[kind] = std::max<int64_t>(room[kind], 0); tot[OPTB_C_STATUS] |= kind ? OPTB_ST_HIT_OVERFLOW : OPTB_ST_SEG_OVERFLOW; }
    for (auto& c : cols) {
      const int64_t r = rows[c.kind], off = c.kind ? hit_off : seg_off;
      if (r > 0) CK(cudaMemcpyAsync(c.host + (size_t)off * c.elt, sorted ? c.srt[k & 1] : c.dev[k & 1], (size_t)r * c.elt, cudaMemcpyDeviceToHost, ctx->s_d2h), "D2H rows");
    }
    CK(cudaEventRecord(ev_d2h[k], ctx->s_d2h), "event");
    if (prm->record_segments) seg_off += rows[0];
    if (prm->record_hits) hit_off += rows[1];
    for (int q : {OPTB_C_SEGMENTS, OPTB_C_INTERACTIONS, OPTB_C_HITS, OPTB_C_TESTS, OPTB_C_DROPPED, OPTB_C_LAUNCHES}) tot[q] += hc[q];
    tot[OPTB_C_STATUS] |= hc[OPTB_C_STATUS];
    tot[OPTB_C_GENERATIONS] = std::max(tot[OPTB_C_GENERATIONS], hc[OPTB_C_GENERATIONS]);
    return 0;
  };
  for (int c = 0; c < nch && !rc; c++) {
    const int64_t lo = (int64_t)c * kChunk, m = std::min<int64_t>(kChunk, n - lo);
    for (int f = 0; f < kRayF64; f++) {
      if (!src_f[f] || ((rays->broadcast >> f) & 1u)) continue;
      CK(cudaMemcpyAsync((void*)(dst_f[f] + lo), src_f[f] + lo, (size_t)m * 8, cudaMemcpyHostToDevice, ctx->s_h2d), "H2D rays");
    }
    if (rays->flags) CK(cudaMemcpyAsync((void*)(dr.flags + lo), rays->flags + lo, (size_t)m * 4, cudaMemcpyHostToDevice, ctx->s_h2d), "H2D flags");
    if (rays->family) CK(cudaMemcpyAsync((void*)(dr.family + lo), rays->family + lo, (size_t)m * 4, cudaMemcpyHostToDevice, ctx->s_h2d), "H2D family");
    CK(cudaEventRecord(ev_h2d[c], ctx->s_h2d), "event");
    CK(cudaStreamWaitEvent(ctx->s_run, ev_h2d[c], 0), "wait");
    if (c >= 2 && !sorted) CK(cudaStreamWaitEvent(ctx->s_run, ev_d2h[c - 2], 0), "wait");  // the result set is free again
    optb_rays cr = dr;
    cr.n = m;
    const double** cf = &cr.ox;
    for (int f = 0; f < kRayF64; f++) if (cf[f] && !((rays->broadcast >> f) & 1u)) cf[f] += lo;
    if (cr.flags) cr.flags += lo;
    if (cr.family) cr.family += lo;
    rc = trace_impl(ctx, scene, &cr, prm, &dv[c & 1], ws[c & 1], wsb, ctx->s_run, (uint32_t)lo, false);
    if (rc) break;
    if (sorted) {
      // reference order inside the chunk (chunks follow each other in ray order): keys + radix sort + one gather of
      // all columns into the sorted set, on the trace stream right behind the trace. The row counts stay on the
      // device (rows beyond them sort to the end), so nothing here waits for the host.
      if (c >= 2) CK(cudaStreamWaitEvent(ctx->s_run, ev_d2h[c - 2], 0), "wait");  // the sorted set is free again
      rc = sort_rows_dev(ctx, &dv[c & 1], &sv[c & 1], prm->record_segments ? cseg : 0, prm->record_hits ? chit : 0,
                         (const unsigned long long*)dv[c & 1].counters, sort_ws, sortb, ctx->s_run, end_bit);
      if (rc) break;
    }
    CK(cudaMemcpyAsync(ctx->h_chunk + (size_t)c * OPTB_C_COUNT, dv[c & 1].counters, OPTB_C_COUNT * 8, cudaMemcpyDeviceToHost, ctx->s_run), "D2H counters");
    CK(cudaEventRecord(ev_run[c], ctx->s_run), "event");
    if (c >= 1) rc = drain(c - 1);
  }
  if (!rc) rc = drain(nch - 1);
  if (!rc && prm->record_hist && scene->n_mons) {
    cudaStreamWaitEvent(ctx->s_d2h, ev_run[nch - 1], 0);
    if (out->hist_y) cudaMemcpyAsync(out->hist_y, d_hy, (size_t)scene->n_mons * OPTB_HIST_BINS * 8, cudaMemcpyDeviceToHost, ctx->s_d2h);
    if (out->hist_yz) cudaMemcpyAsync(out->hist_yz, d_hyz, (size_t)scene->n_mons * OPTB_HIST_BINS * OPTB_HIST_BINS * 8, cudaMemcpyDeviceToHost, ctx->s_d2h);
  }
  cudaStreamSynchronize(ctx->s_h2d); cudaStreamSynchronize(ctx->s_run);
  cudaError_t e = cudaStreamSynchronize(ctx->s_d2h);
  if (rc) return rc;
  if (e != cudaSuccess) return fail(ctx, -10, "pipelined trace", e);
  if (chunk_overflow) return 1;
  memcpy(ctx->h_counters, tot, sizeof tot);
  if (out->counters) memcpy(out->counters, tot, sizeof tot);
  return 0;
}

extern "C" int optb_trace_host(optb_ctx* ctx, const optb_scene* scene, const optb_rays* rays, const optb_params* prm,
                               optb_result* out) {
  if (!ctx || !scene || !rays || !prm || !out) return -1;
  if (!needs_wavefront(scene, prm) && scene->n_caps == 0 && rays->n >= (3ll << 20)) {
    int rc = trace_host_pipelined(ctx, scene, rays, prm, out);
    if (rc <= 0) return rc;  // 1 = a chunk outgrew its estimated share of the result capacity: one-shot path
  }
  // The live ray set of a splitting scene is not known in advance: grow the workspace until it fits. A root pops at
  // most max_trace_num rays and every pop queues at most two, so 2 n max_trace_num live rays always suffice.
  const double bound_d = 2.0 * (double)std::max<int64_t>(rays->n, 1) * (double)std::max<int64_t>(prm->max_trace_num, 1);
  const int64_t bound = bound_d > 4e9 ? (int64_t)4e9 : (int64_t)bound_d;
  const int64_t per_ray = std::max<int64_t>(4, ctx->live_per_ray_hint);
  for (int64_t live = std::min<int64_t>(std::max<int64_t>(per_ray * rays->n, 1024), std::max<int64_t>(bound, 1024));;
       live = std::min<int64_t>(live * 2, bound)) {
    int rc = trace_host_once(ctx, scene, rays, prm, out, live);
    if (rc) return rc;
    if (!(ctx->h_counters[OPTB_C_STATUS] & OPTB_ST_WORK_OVERFLOW)) {
      if (rays->n > 0 && needs_wavefront(scene, prm)) ctx->live_per_ray_hint = (live + rays->n - 1) / rays->n;
      return 0;
    }
    if (live >= bound) break;
  }
  return fail(ctx, -8, "live ray set exceeds 2 * n * max_trace_num: workspace overflow");
}

static int trace_host_once(optb_ctx* ctx, const optb_scene* scene, const optb_rays* rays, const optb_params* prm,
                           optb_result* out, int64_t live) {
  cudaSetDevice(ctx->device);
  const int64_t n = rays->n;
  const bool split = needs_wavefront(scene, prm) || needs_serial(scene, rays, prm);
  const int64_t segcap = prm->record_segments ? out->seg_capacity : 0, hitcap = prm->record_hits ? out->hit_capacity : 0;
  const int64_t max_live = split ? std::max<int64_t>(live, n) : 0;
  const int64_t wsb = needs_serial(scene, rays, prm) ? optb_workspace_bytes(scene, n, 16 * std::max<int64_t>(live, 64))
                                               : (int64_t)ws_layout(n, max_live, split).total;
  const int nfam = std::max(prm->n_families, 1);
  const int64_t sortb = prm->sorted_rows ? optb_sort_workspace_bytes(std::max(segcap, hitcap)) : 0;
  size_t need = 256 * 64 + (size_t)sortb + (size_t)n * (8 * 13 + 8) + (size_t)segcap * (13 * 8 + 16) + (size_t)hitcap * (10 * 8 + 12 + 8) +
                (size_t)std::max(scene->n_mons, 1) * OPTB_HIST_BINS * (OPTB_HIST_BINS + 1) * 8 +
                (size_t)std::max(scene->n_caps, 1) * nfam * 4 + 64 * 8 + (size_t)wsb;
  if (ctx->arena_bytes < need) {
    if (ctx->arena) cudaFree(ctx->arena);
    ctx->arena = nullptr; ctx->arena_bytes = 0;
    CK(cudaMalloc(&ctx->arena, need), "cudaMalloc(arena)");
    ctx->arena_bytes = need;
  }
  ArenaCursor ac{(unsigned char*)ctx->arena, 0, ctx->arena_bytes};
  cudaStream_t st = 0;
  optb_rays dr = *rays;
  const double* const* src_f = &rays->ox;
  const double** dst_f = &dr.ox;
  for (int f = 0; f < kRayF64; f++) {
    if (!src_f[f]) { dst_f[f] = nullptr; continue; }
    const size_t rows = ((rays->broadcast >> f) & 1u) ? 1 : (size_t)n;
    void* d = ac.take(rows * 8);
    if (!d) return fail(ctx, -9, "arena sizing");
    CK(cudaMemcpyAsync(d, src_f[f], rows * 8, cudaMemcpyHostToDevice, st), "H2D rays");
    dst_f[f] = (const double*)d;
  }
  if (rays->flags) { void* d = ac.take((size_t)n * 4); CK(cudaMemcpyAsync(d, rays->flags, (size_t)n * 4, cudaMemcpyHostToDevice, st), "H2D flags"); dr.flags = (const uint32_t*)d; }
  if (rays->family) { void* d = ac.take((size_t)n * 4); CK(cudaMemcpyAsync(d, rays->family, (size_t)n * 4, cudaMemcpyHostToDevice, st), "H2D family"); dr.family = (const int32_t*)d; }
  optb_result dv = *out;
  struct Col { void** dev; void* host; size_t elt; int64_t cap; int kind; };  // kind 0 seg, 1 hit
  std::vector<Col> cols;
  auto add = [&](void** dev_field, void* host_ptr, size_t elt, int64_t capn, int kind) {
    if (!host_ptr || capn == 0) { *dev_field = nullptr; return; }
    *dev_field = ac.take((size_t)capn * elt);
    cols.push_back({dev_field, host_ptr, elt, capn, kind});
  };
  double** seg_d = &dv.seg_ox; double* const* seg_h = &out->seg_ox;
  for (int f = 0; f < 13; f++) add((void**)&seg_d[f], seg_h[f], 8, segcap, 0);
  add((void**)&dv.seg_flags, out->seg_flags, 4, segcap, 0); add((void**)&dv.seg_root, out->seg_root, 4, segcap, 0);
  add((void**)&dv.seg_pop, out->seg_pop, 4, segcap, 0); add((void**)&dv.seg_leaf, out->seg_leaf, 4, segcap, 0);
  add((void**)&dv.hit_monitor, out->hit_monitor, 4, hitcap, 1); add((void**)&dv.hit_root, out->hit_root, 4, hitcap, 1);
  add((void**)&dv.hit_pop, out->hit_pop, 4, hitcap, 1);
  double** hit_d = &dv.hit_px; double* const* hit_h = &out->hit_px;
  for (int f = 0; f < 10; f++) add((void**)&hit_d[f], hit_h[f], 8, hitcap, 1);
  add((void**)&dv.hit_key, out->hit_key, 8, hitcap, 1);
  dv.root_flags = nullptr;
  for (auto& c : cols) if (!*c.dev) return fail(ctx, -9, "arena sizing");
  void* sort_ws = sortb ? ac.take((size_t)sortb) : nullptr;
  if (sortb && !sort_ws) return fail(ctx, -9, "arena sizing");
  const size_t hy = (size_t)std::max(scene->n_mons, 1) * OPTB_HIST_BINS * 8, hyz = hy * OPTB_HIST_BINS;
  dv.hist_y = (int64_t*)ac.take(hy); dv.hist_yz = (int64_t*)ac.take(hyz);
  const size_t capb = (size_t)std::max(scene->n_caps, 1) * nfam * 4;
  dv.cap_counts = (int32_t*)ac.take(capb);
  dv.counters = (int64_t*)ac.take(OPTB_C_COUNT * 8);
  void* ws = ac.take((size_t)wsb);
  if (!ws || !dv.counters) return fail(ctx, -9, "arena sizing");
  if (scene->n_caps > 0 && out->cap_counts) CK(cudaMemcpyAsync(dv.cap_counts, out->cap_counts, capb, cudaMemcpyHostToDevice, st), "H2D caps");
  else CK(cudaMemsetAsync(dv.cap_counts, 0, capb, st), "memset caps");  // no table given: every family starts at zero
  int rc = optb_trace(ctx, scene, &dr, prm, &dv, ws, wsb, st);
  if (rc) return rc;
  CK(cudaMemcpyAsync(ctx->h_counters, dv.counters, OPTB_C_COUNT * 8, cudaMemcpyDeviceToHost, st), "D2H counters");
  CK(cudaStreamSynchronize(st), "sync trace");
  if (out->counters) memcpy(out->counters, ctx->h_counters, OPTB_C_COUNT * 8);
  int64_t nseg = std::min<int64_t>((int64_t)ctx->h_counters[OPTB_C_SEGMENTS], segcap);
  int64_t nhit = std::min<int64_t>((int64_t)ctx->h_counters[OPTB_C_HITS], hitcap);
  if (prm->sorted_rows)
    if (int src = optb_sort_rows(ctx, &dv, prm->record_segments ? nseg : 0, prm->record_hits ? nhit : 0, sort_ws, sortb, st)) return src;
  for (auto& c : cols) {
    int64_t rows = c.kind == 0 ? nseg : nhit;
    if (rows > 0) CK(cudaMemcpyAsync(c.host, *c.dev, (size_t)rows * c.elt, cudaMemcpyDeviceToHost, st), "D2H results");
  }
  if (prm->record_hist && scene->n_mons) {
    if (out->hist_y) CK(cudaMemcpyAsync(out->hist_y, dv.hist_y, (size_t)scene->n_mons * OPTB_HIST_BINS * 8, cudaMemcpyDeviceToHost, st), "D2H hist");
    if (out->hist_yz) CK(cudaMemcpyAsync(out->hist_yz, dv.hist_yz, (size_t)scene->n_mons * OPTB_HIST_BINS * OPTB_HIST_BINS * 8, cudaMemcpyDeviceToHost, st), "D2H hist");
  }
  if (scene->n_caps > 0 && out->cap_counts) CK(cudaMemcpyAsync(out->cap_counts, dv.cap_counts, capb, cudaMemcpyDeviceToHost, st), "D2H caps");
  CK(cudaStreamSynchronize(st), "sync results");
  return 0;
}

extern "C" int optb_measure_fp64_peak(optb_ctx* ctx, double* tflops_out) {
  if (!ctx || !tflops_out) return -1;
  cudaSetDevice(ctx->device);
  const int blocks = ctx->sm_count * 8, threads = 256, iters = 1 << 16;
  double* d = nullptr;
  CK(cudaMalloc((void**)&d, sizeof(double) * blocks * threads), "cudaMalloc");
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  dfma_kernel<<<blocks, threads>>>(d, 1024);
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    dfma_kernel<<<blocks, threads>>>(d, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    best = std::min(best, ms);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaError_t e = cudaGetLastError();
  cudaFree(d);
  if (e != cudaSuccess) return fail(ctx, -10, "dfma kernel", e);
  double flops = 2.0 * 8.0 * (double)iters * blocks * threads;
  *tflops_out = flops / (best * 1e-3) / 1e12;
  return 0;
}


// ---- rows into reference order ------------------------------------------------------------------------------------
namespace {
// `d_count` (optional): the number of valid rows lives on the device (a counter of the trace that just ran on the same
// stream); rows beyond it get the largest key and sort to the end, so the host need not know the count to enqueue the sort
__global__ void seg_keys_kernel(const uint32_t* __restrict__ root, const uint32_t* __restrict__ pop, long long n,
                                const unsigned long long* __restrict__ d_count,
                                unsigned long long* __restrict__ key, uint32_t* __restrict__ idx) {
  const long long valid = d_count ? (long long)min(*d_count, (unsigned long long)n) : n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    key[i] = i < valid ? (((unsigned long long)root[i] << 32) | pop[i]) : ~0ull;
    idx[i] = (uint32_t)i;
  }
}
__global__ void hit_keys_kernel(const unsigned long long* __restrict__ packed, const uint32_t* __restrict__ root,
                                const int32_t* __restrict__ mon, const uint32_t* __restrict__ pop, long long n,
                                const unsigned long long* __restrict__ d_count,
                                unsigned long long* __restrict__ key, uint32_t* __restrict__ idx) {
  const long long valid = d_count ? (long long)min(*d_count, (unsigned long long)n) : n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    key[i] = i >= valid ? ~0ull
             : packed   ? packed[i]
                        : (((unsigned long long)root[i] << 32) | ((unsigned long long)(uint32_t)mon[i] << 24) | pop[i]);
    idx[i] = (uint32_t)i;
  }
}
// all columns of a row set in one pass: dst[c][i] = src[c][order[i]]
struct GatherCols { const void* src[20]; void* dst[20]; int elt[20]; int ncol; };
__global__ void __launch_bounds__(256) gather_rows_kernel(const GatherCols g, const uint32_t* __restrict__ order, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const uint32_t j = order[i];
    for (int c = 0; c < g.ncol; c++) {
      if (g.elt[c] == 8) ((unsigned long long*)g.dst[c])[i] = ((const unsigned long long*)g.src[c])[j];
      else ((uint32_t*)g.dst[c])[i] = ((const uint32_t*)g.src[c])[j];
    }
  }
}
template <class T>
__global__ void gather_kernel(const T* __restrict__ src, const uint32_t* __restrict__ idx, long long n, T* __restrict__ dst) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = src[idx[i]];
}
struct SortLayout { size_t key_a, key_b, idx_a, idx_b, tmp, cub, cub_bytes, total; };
SortLayout sort_layout(long long n) {
  SortLayout L{};
  n = std::max<long long>(n, 1);
  size_t o = 0;
  L.key_a = o; o += align_up((size_t)n * 8, 256);
  L.key_b = o; o += align_up((size_t)n * 8, 256);
  L.idx_a = o; o += align_up((size_t)n * 4, 256);
  L.idx_b = o; o += align_up((size_t)n * 4, 256);
  L.tmp = o; o += align_up((size_t)n * 8, 256);
  size_t tb = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tb, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                  (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)std::min<long long>(n, 0x7fffffffll), 0, 64);
  L.cub = o; L.cub_bytes = align_up(tb + 256, 256); o += L.cub_bytes;
  L.total = o;
  return L;
}
template <class T>
int reorder_column(optb_ctx* ctx, T* col, const uint32_t* order, long long n, void* tmp, int grid, cudaStream_t st) {
  if (!col) return 0;
  gather_kernel<T><<<grid, 256, 0, st>>>(col, order, n, (T*)tmp);
  CK(cudaMemcpyAsync(col, tmp, (size_t)n * sizeof(T), cudaMemcpyDeviceToDevice, st), "reorder column");
  return 0;
}
}  // namespace

extern "C" int64_t optb_sort_workspace_bytes(int64_t n_rows) { return n_rows < 0 ? -1 : (int64_t)sort_layout(n_rows).total; }

// Sort the first n_seg / n_hit rows of `src` by their keys. dst == nullptr: in place (column by column through a
// scratch column); else every present column of src is gathered into the same column of dst in one pass.
// d_counters (optional): device counters of the trace that produced the rows; then n_seg / n_hit are capacities.
static int sort_rows_dev(optb_ctx* ctx, const optb_result* res, const optb_result* dst, int64_t n_seg, int64_t n_hit,
                         const unsigned long long* d_counters, void* workspace, int64_t workspace_bytes, cudaStream_t st,
                         int end_bit) {
  const long long nmax = std::max<long long>(n_seg, n_hit);
  if (nmax <= 0) return 0;
  if (nmax >= 0x7fffffffll) return fail(ctx, -7, "optb_sort_rows: at most 2^31-1 rows");
  const SortLayout L = sort_layout(nmax);
  if (!workspace || workspace_bytes < (int64_t)L.total) return fail(ctx, -8, "optb_sort_rows: workspace too small (optb_sort_workspace_bytes)");
  unsigned char* ws = (unsigned char*)workspace;
  unsigned long long* key_a = (unsigned long long*)(ws + L.key_a);
  unsigned long long* key_b = (unsigned long long*)(ws + L.key_b);
  uint32_t* idx_a = (uint32_t*)(ws + L.idx_a);
  uint32_t* idx_b = (uint32_t*)(ws + L.idx_b);
  void* tmp = ws + L.tmp;
  for (int pass = 0; pass < 2; pass++) {
    const long long n = pass == 0 ? n_seg : n_hit;
    if (n <= 1 && !d_counters) continue;
    if (n <= 0) continue;
    const int grid = (int)std::min<long long>((n + 255) / 256, (long long)ctx->sm_count * 16);
    if (pass == 0) {
      if (!res->seg_root || !res->seg_pop) return fail(ctx, -7, "optb_sort_rows: seg_root and seg_pop are the segment key");
      seg_keys_kernel<<<grid, 256, 0, st>>>(res->seg_root, res->seg_pop, n, d_counters ? d_counters + OPTB_C_SEGMENTS : nullptr, key_a, idx_a);
    } else {
      if (!res->hit_key && !(res->hit_root && res->hit_pop && res->hit_monitor))
        return fail(ctx, -7, "optb_sort_rows: hit_key, or hit_root + hit_monitor + hit_pop, is the monitor-row key");
      hit_keys_kernel<<<grid, 256, 0, st>>>((const unsigned long long*)res->hit_key, res->hit_root, res->hit_monitor, res->hit_pop, n,
                                            d_counters ? d_counters + OPTB_C_HITS : nullptr, key_a, idx_a);
    }
    size_t tb = L.cub_bytes;
    CK(cub::DeviceRadixSort::SortPairs(ws + L.cub, tb, (const unsigned long long*)key_a, key_b, (const uint32_t*)idx_a, idx_b,
                                       (int)n, 0, end_bit, st), "radix sort of the row keys");
    if (dst) {
      GatherCols g;
      g.ncol = 0;
      auto col = [&](const void* sp, void* dp, int elt) { if (sp && dp) { g.src[g.ncol] = sp; g.dst[g.ncol] = dp; g.elt[g.ncol] = elt; g.ncol++; } };
      if (pass == 0) {
        double* const* f = &res->seg_ox; double* const* fd = &dst->seg_ox;
        for (int c = 0; c < 13; c++) col(f[c], fd[c], 8);
        col(res->seg_flags, dst->seg_flags, 4); col(res->seg_root, dst->seg_root, 4);
        col(res->seg_pop, dst->seg_pop, 4); col(res->seg_leaf, dst->seg_leaf, 4);
      } else {
        double* const* f = &res->hit_px; double* const* fd = &dst->hit_px;
        for (int c = 0; c < 10; c++) col(f[c], fd[c], 8);
        col(res->hit_monitor, dst->hit_monitor, 4); col(res->hit_root, dst->hit_root, 4);
        col(res->hit_pop, dst->hit_pop, 4); col(res->hit_key, dst->hit_key, 8);
      }
      if (g.ncol) gather_rows_kernel<<<grid, 256, 0, st>>>(g, idx_b, n);
      continue;
    }
    int rc = 0;
    if (pass == 0) {
      double* const* f = &res->seg_ox;
      for (int c = 0; c < 13 && !rc; c++) rc = reorder_column<double>(ctx, f[c], idx_b, n, tmp, grid, st);
      if (!rc) rc = reorder_column<uint32_t>(ctx, res->seg_flags, idx_b, n, tmp, grid, st);
      if (!rc) rc = reorder_column<uint32_t>(ctx, res->seg_root, idx_b, n, tmp, grid, st);
      if (!rc) rc = reorder_column<uint32_t>(ctx, res->seg_pop, idx_b, n, tmp, grid, st);
      if (!rc) rc = reorder_column<int32_t>(ctx, res->seg_leaf, idx_b, n, tmp, grid, st);
    } else {
      double* const* f = &res->hit_px;
      for (int c = 0; c < 10 && !rc; c++) rc = reorder_column<double>(ctx, f[c], idx_b, n, tmp, grid, st);
      if (!rc) rc = reorder_column<int32_t>(ctx, res->hit_monitor, idx_b, n, tmp, grid, st);
      if (!rc) rc = reorder_column<uint32_t>(ctx, res->hit_root, idx_b, n, tmp, grid, st);
      if (!rc) rc = reorder_column<uint32_t>(ctx, res->hit_pop, idx_b, n, tmp, grid, st);
      if (!rc) rc = reorder_column<unsigned long long>(ctx, (unsigned long long*)res->hit_key, idx_b, n, tmp, grid, st);
    }
    if (rc) return rc;
  }
  CK(cudaGetLastError(), "optb_sort_rows");
  return 0;
}

extern "C" int optb_sort_rows(optb_ctx* ctx, const optb_result* res, int64_t n_seg, int64_t n_hit, void* workspace,
                              int64_t workspace_bytes, void* stream_v) {
  if (!ctx || !res) return -1;
  cudaSetDevice(ctx->device);
  return sort_rows_dev(ctx, res, nullptr, n_seg, n_hit, nullptr, workspace, workspace_bytes, (cudaStream_t)stream_v, 64);
}

// ---- monitor analytics: one fused pass over the rows in HBM ---------------------------------------------------------
namespace {
struct StatsArgs {
  const double *px, *py, *pz, *I, *t, *dx, *dy, *dz, *qre;
  const int32_t* mon; const unsigned long long* key;
  long long first, n; int monitor;
  optb_monitor_frame f;
  double *stats, *y, *z, *ty, *tz, *wd;
};
OPTB_DEV void atomic_min_f64(double* addr, double v) {
  unsigned long long* a = (unsigned long long*)addr;
  unsigned long long old = *a;
  while (v < __longlong_as_double((long long)old)) {
    const unsigned long long seen = atomicCAS(a, old, (unsigned long long)__double_as_longlong(v));
    if (seen == old) break;
    old = seen;
  }
}
OPTB_DEV void atomic_max_f64(double* addr, double v) {
  unsigned long long* a = (unsigned long long*)addr;
  unsigned long long old = *a;
  while (v > __longlong_as_double((long long)old)) {
    const unsigned long long seen = atomicCAS(a, old, (unsigned long long)__double_as_longlong(v));
    if (seen == old) break;
    old = seen;
  }
}
__global__ void stats_init_kernel(double* s) {
  const int k = threadIdx.x;
  if (k < OPTB_MS_STRIDE) s[k] = (k == OPTB_MS_MIN_Y || k == OPTB_MS_MIN_Z) ? INFINITY : (k == OPTB_MS_MAX_Y || k == OPTB_MS_MAX_Z) ? -INFINITY : 0.0;
}
__global__ void __launch_bounds__(256) monitor_stats_kernel(const StatsArgs a) {
  __shared__ unsigned int s_hist[OPTB_HIST_BINS];
  __shared__ double s_red[8][8];
  if (threadIdx.x < OPTB_HIST_BINS) s_hist[threadIdx.x] = 0u;
  __syncthreads();
  double cnt = 0, sI = 0, sy = 0, syy = 0, sz = 0, szz = 0, swd = 0, sty = 0, styty = 0;
  double ymin = INFINITY, ymax = -INFINITY, zmin = INFINITY, zmax = -INFINITY;
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < a.n; k += (long long)gridDim.x * blockDim.x) {
    const long long j = a.first + k;
    bool mine = true;
    if (a.monitor >= 0) mine = a.mon ? (a.mon[j] == a.monitor) : (int)((a.key[j] >> 24) & 0xffu) == a.monitor;
    double y = NAN, z = NAN, ty = NAN, tz = NAN, wd = NAN;
    if (mine) {
      const double Px = a.px[j], Py = a.py[j], Pz = a.pz[j];
      y = dot3(Px, Py, Pz, a.f.tangent_y[0], a.f.tangent_y[1], a.f.tangent_y[2]);
      z = dot3(Px, Py, Pz, a.f.tangent_z[0], a.f.tangent_z[1], a.f.tangent_z[2]);
      cnt += 1.0; sI += a.I[j]; sy += y; syy = fma(y, y, syy); sz += z; szz = fma(z, z, szz);
      ymin = fmin(ymin, y); ymax = fmax(ymax, y); zmin = fmin(zmin, z); zmax = fmax(zmax, z);
      const int b = hist_bin(y, -a.f.half_width, a.f.half_width);
      if (b >= 0) atomicAdd(&s_hist[b], 1u);
      if (a.dx) {
        const double dx = a.dx[j], dy = a.dy[j], dz = a.dz[j];
        ty = dot3(dx, dy, dz, a.f.tangent_y[0], a.f.tangent_y[1], a.f.tangent_y[2]);
        tz = dot3(dx, dy, dz, a.f.tangent_z[0], a.f.tangent_z[1], a.f.tangent_z[2]);
        sty += ty; styty = fma(ty, ty, styty);
        if (a.qre) {
          const double dist = a.qre[j] + a.t[j];  // Re(q_at_z(t)) = distance to the waist plane (ray.py:17-25)
          wd = dot3(dx, dy, dz, a.f.normal[0], a.f.normal[1], a.f.normal[2]) > 0.0 ? -dist : dist;
          swd += wd;
        }
      }
    }
    if (a.y) a.y[k] = y;
    if (a.z) a.z[k] = z;
    if (a.ty) a.ty[k] = ty;
    if (a.tz) a.tz[k] = tz;
    if (a.wd) a.wd[k] = wd;
  }
  // block reduction: warp shuffles, then one atomic per block and quantity
  double v[9] = {cnt, sI, sy, syy, sz, szz, swd, sty, styty};
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 9; q++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ymin = fmin(ymin, __shfl_xor_sync(0xffffffffu, ymin, o)); ymax = fmax(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
    zmin = fmin(zmin, __shfl_xor_sync(0xffffffffu, zmin, o)); zmax = fmax(zmax, __shfl_xor_sync(0xffffffffu, zmax, o));
  }
  const int slot[9] = {OPTB_MS_COUNT, OPTB_MS_SUM_I, OPTB_MS_SUM_Y, OPTB_MS_SUM_YY, OPTB_MS_SUM_Z, OPTB_MS_SUM_ZZ,
                       OPTB_MS_SUM_WD, OPTB_MS_SUM_TY, OPTB_MS_SUM_TYTY};
  if (lane == 0) {
    for (int q = 0; q < 8; q++) s_red[q][wid] = q < 8 ? v[q] : 0.0;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    double tot = 0;
    for (int w = 0; w < 8; w++) tot += s_red[threadIdx.x][w];
    if (tot != 0.0) atomicAdd(&a.stats[slot[threadIdx.x]], tot);
  }
  if (lane == 0) {
    if (v[8] != 0.0) atomicAdd(&a.stats[OPTB_MS_SUM_TYTY], v[8]);
    if (ymin <= ymax) { atomic_min_f64(&a.stats[OPTB_MS_MIN_Y], ymin); atomic_max_f64(&a.stats[OPTB_MS_MAX_Y], ymax); }
    if (zmin <= zmax) { atomic_min_f64(&a.stats[OPTB_MS_MIN_Z], zmin); atomic_max_f64(&a.stats[OPTB_MS_MAX_Z], zmax); }
  }
  __syncthreads();
  if (threadIdx.x < OPTB_HIST_BINS && s_hist[threadIdx.x]) atomicAdd(&a.stats[OPTB_MS_HIST + threadIdx.x], (double)s_hist[threadIdx.x]);
}
}  // namespace

extern "C" int optb_monitor_stats(optb_ctx* ctx, const optb_result* rows, int64_t first, int64_t n, int monitor,
                                  const optb_monitor_frame* frame, double* stats, double* y, double* z, double* ty,
                                  double* tz, double* waist, void* stream_v) {
  if (!ctx || !rows || !frame || !stats || first < 0 || n < 0) return -1;
  if (!rows->hit_px || !rows->hit_py || !rows->hit_pz || !rows->hit_intensity || !rows->hit_t)
    return fail(ctx, -7, "optb_monitor_stats: hit_px/py/pz, hit_intensity and hit_t are required");
  if (monitor >= 0 && !rows->hit_monitor && !rows->hit_key) return fail(ctx, -7, "optb_monitor_stats: selecting a monitor needs hit_monitor or hit_key");
  if ((ty || tz || waist) && !(rows->hit_dx && rows->hit_dy && rows->hit_dz)) return fail(ctx, -7, "optb_monitor_stats: slopes / waist distances need hit_dx, hit_dy, hit_dz");
  if (waist && !rows->hit_q_re) return fail(ctx, -7, "optb_monitor_stats: waist distances need hit_q_re");
  cudaSetDevice(ctx->device);
  cudaStream_t st = (cudaStream_t)stream_v;
  StatsArgs a;
  a.px = rows->hit_px; a.py = rows->hit_py; a.pz = rows->hit_pz; a.I = rows->hit_intensity; a.t = rows->hit_t;
  a.dx = rows->hit_dx && rows->hit_dy && rows->hit_dz ? rows->hit_dx : nullptr; a.dy = rows->hit_dy; a.dz = rows->hit_dz;
  a.qre = rows->hit_q_re; a.mon = rows->hit_monitor; a.key = (const unsigned long long*)rows->hit_key;
  a.first = first; a.n = n; a.monitor = monitor; a.f = *frame;
  a.stats = stats; a.y = y; a.z = z; a.ty = ty; a.tz = tz; a.wd = waist;
  stats_init_kernel<<<1, 64, 0, st>>>(stats);
  if (n > 0) {
    const int grid = (int)std::min<long long>((n + 255) / 256, (long long)ctx->sm_count * 8);
    monitor_stats_kernel<<<grid, 256, 0, st>>>(a);
  }
  CK(cudaGetLastError(), "monitor_stats_kernel");
  return 0;
}

// ---- multi-GPU monitor merge over NCCL (loaded at run time: no link-time dependency) ------------------------------
namespace {
// the few NCCL entry points used, with the ABI of nccl.h (ncclResult_t = int, ncclUniqueId = 128 bytes by value)
struct NcclId { char internal[128]; };
struct NcclApi {
  int (*GetUniqueId)(NcclId*);
  int (*CommInitRank)(void**, int, NcclId, int);
  int (*CommDestroy)(void*);
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
  int (*GroupStart)();
  int (*GroupEnd)();
  const char* (*GetErrorString)(int);
};
constexpr int kNcclInt64 = 4, kNcclSum = 0, kNcclMax = 2;  // ncclDataType_t / ncclRedOp_t values of nccl.h
NcclApi g_nccl;
void* g_nccl_handle = nullptr;

int load_nccl(optb_ctx* ctx) {
  if (g_nccl_handle) return 0;
  void* h = nullptr;
  if (const char* p = getenv("OPTB_NCCL_LIB")) h = dlopen(p, RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // already in the process (e.g. torch's)
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return fail(ctx, -11, "NCCL not found: set OPTB_NCCL_LIB to libnccl.so.2");
  NcclApi a;
  a.GetUniqueId = (int (*)(NcclId*))dlsym(h, "ncclGetUniqueId");
  a.CommInitRank = (int (*)(void**, int, NcclId, int))dlsym(h, "ncclCommInitRank");
  a.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
  a.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclAllReduce");
  a.GroupStart = (int (*)())dlsym(h, "ncclGroupStart");
  a.GroupEnd = (int (*)())dlsym(h, "ncclGroupEnd");
  a.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
  if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllReduce || !a.GroupStart || !a.GroupEnd)
    return fail(ctx, -11, "NCCL library lacks an expected symbol");
  g_nccl = a; g_nccl_handle = h;
  return 0;
}
int nccl_fail(optb_ctx* ctx, const char* what, int rc) {
  snprintf(ctx->err, sizeof ctx->err, "%s: NCCL error %d (%s)", what, rc, g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
  return -11;
}

// status word -> one slot per bit (so that a sum over ranks is an OR), generations -> their own slot (maximum)
constexpr int kMergeBits = 16;
__global__ void merge_prep_kernel(long long* counters, long long* scratch) {
  const int k = threadIdx.x;
  if (k < kMergeBits) scratch[k] = (counters[OPTB_C_STATUS] >> k) & 1;
  if (k == kMergeBits) scratch[kMergeBits] = counters[OPTB_C_GENERATIONS];
  __syncthreads();
  if (k == 0) { counters[OPTB_C_STATUS] = 0; counters[OPTB_C_GENERATIONS] = 0; }
}
__global__ void merge_post_kernel(long long* counters, const long long* scratch) {
  if (threadIdx.x == 0) {
    long long st = 0;
    for (int k = 0; k < kMergeBits; k++) if (scratch[k] > 0) st |= 1ll << k;
    counters[OPTB_C_STATUS] = st;
    counters[OPTB_C_GENERATIONS] = scratch[kMergeBits];
  }
}
}  // namespace

extern "C" int optb_comm_unique_id(optb_ctx* ctx, void* id128) {
  if (!ctx || !id128) return -1;
  if (int rc = load_nccl(ctx)) return rc;
  NcclId id;
  if (int rc = g_nccl.GetUniqueId(&id)) return nccl_fail(ctx, "ncclGetUniqueId", rc);
  memcpy(id128, &id, sizeof id);
  return 0;
}

extern "C" int optb_comm_init(optb_ctx* ctx, const void* id128, int rank, int nranks) {
  if (!ctx || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return -1;
  if (ctx->nccl_comm) return fail(ctx, -7, "communicator already initialised");
  if (int rc = load_nccl(ctx)) return rc;
  cudaSetDevice(ctx->device);
  NcclId id;
  memcpy(&id, id128, sizeof id);
  void* comm = nullptr;
  if (int rc = g_nccl.CommInitRank(&comm, nranks, id, rank)) return nccl_fail(ctx, "ncclCommInitRank", rc);
  CK(cudaMalloc((void**)&ctx->d_merge, sizeof(long long) * (kMergeBits + 1)), "cudaMalloc(merge scratch)");
  ctx->nccl_comm = comm; ctx->comm_rank = rank; ctx->comm_size = nranks;
  return 0;
}

extern "C" int optb_monitor_merge(optb_ctx* ctx, int64_t* hist_y, int64_t* hist_yz, int n_monitors, int64_t* counters,
                                  void* stream_v) {
  if (!ctx || n_monitors < 0) return -1;
  if (!ctx->nccl_comm) return fail(ctx, -7, "optb_monitor_merge: call optb_comm_init first");
  cudaSetDevice(ctx->device);
  cudaStream_t st = (cudaStream_t)stream_v;
  if (ctx->comm_size == 1) return 0;
  if (counters) merge_prep_kernel<<<1, 32, 0, st>>>((long long*)counters, ctx->d_merge);
  int rc = g_nccl.GroupStart();
  if (!rc && hist_y && n_monitors) rc = g_nccl.AllReduce(hist_y, hist_y, (size_t)n_monitors * OPTB_HIST_BINS, kNcclInt64, kNcclSum, ctx->nccl_comm, st);
  if (!rc && hist_yz && n_monitors) rc = g_nccl.AllReduce(hist_yz, hist_yz, (size_t)n_monitors * OPTB_HIST_BINS * OPTB_HIST_BINS, kNcclInt64, kNcclSum, ctx->nccl_comm, st);
  if (!rc && counters) rc = g_nccl.AllReduce(counters, counters, OPTB_C_COUNT, kNcclInt64, kNcclSum, ctx->nccl_comm, st);
  if (!rc && counters) rc = g_nccl.AllReduce(ctx->d_merge, ctx->d_merge, kMergeBits, kNcclInt64, kNcclSum, ctx->nccl_comm, st);
  if (!rc && counters) rc = g_nccl.AllReduce(ctx->d_merge + kMergeBits, ctx->d_merge + kMergeBits, 1, kNcclInt64, kNcclMax, ctx->nccl_comm, st);
  int rc2 = g_nccl.GroupEnd();
  if (rc || rc2) return nccl_fail(ctx, "ncclAllReduce", rc ? rc : rc2);
  if (counters) merge_post_kernel<<<1, 32, 0, st>>>((long long*)counters, ctx->d_merge);
  CK(cudaGetLastError(), "merge kernels");
  return 0;
}

extern "C" int optb_comm_destroy(optb_ctx* ctx) {
  if (!ctx) return -1;
  if (ctx->nccl_comm) {
    cudaSetDevice(ctx->device);
    g_nccl.CommDestroy(ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
  }
  if (ctx->d_merge) { cudaFree(ctx->d_merge); ctx->d_merge = nullptr; }
  return 0;
}
