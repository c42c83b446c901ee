// CPU single-stepping of the __host__ __device__ logic in tess2_b200/csrc/cell_core.cuh and
// host_geom.hpp -- TEST ONLY.  It exists because the development container has no GPU: the same
// functions the kernels call (star walk, edge circulation, Newell normals, plane tests, the
// scan-line state machine, CIC weights and the per-grid-point gather of DENSE_CIC, span emission, the accumulate step) are driven here in
// the kernels' order so that logic errors show up against the oracle before a GPU run.
// This file is never linked into libtess_b200.so and is not a fallback: the product has none.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../tess2_b200/csrc/host_geom.hpp"

using namespace tb;

extern "C" {
struct emu_block_t
{
  int gid, num_orig_particles, num_particles;
  const float *particles;
  int num_tets;
  const int *tets;
  const int *vert_to_tet;
  float bounds_min[3], bounds_max[3];
  float *density;
  long long density_capacity;
  int block_min_idx[3], block_num_idx[3];
  int num_grid_pts;
};
struct emu_params_t
{
  int alg, num_given_bounds;
  float given_mins[3], given_maxs[3];
  int project;
  float proj_plane[3];
  float mass, eps;
  int glo_num_idx[3];
  float data_mins[3], data_maxs[3], grid_phys_mins[3], grid_phys_maxs[3], grid_step_size[3];
  double seconds;
};
}

struct HostWS
{
  std::vector<int> s, u, t;
  HostWS() : s(4096), u(1024), t(1024) {}
  int &star(int i) { return s[i]; }
  int &nu(int i) { return u[i]; }
  int &nt(int i) { return t[i]; }
};

// the fast kernels' workspace (hash tables, caps 56 / 36) with the linear-search workspace as the
// overflow path -- the same two-level scheme as k_cell_topo / k_cell_topo_big
struct HashWS
{
  int s[52], u[36];
  unsigned char nti[36], par[52], vh[64], nh[64];
  unsigned char &parent_idx(int i) { return par[i]; }
  int &star(int i) { return s[i]; }
  int &nu(int i) { return u[i]; }
  unsigned char &nt_idx(int i) { return nti[i]; }
  unsigned char &vis_hash(unsigned h) { return vh[h]; }
  unsigned char &nbr_hash(unsigned h) { return nh[h]; }
  int nt(int i) { return s[nti[i]]; }
  void hash_clear() { memset(vh, 0xFF, 64); memset(nh, 0xFF, 64); }
};

// the split fast path of k_cell_bfs + k_cell_nbrs
struct StarHostWS
{
  int s[52];
  unsigned char par[52];
  uint32_t vw[16];
  int &star(int i) { return s[i]; }
  unsigned char &parent_idx(int i) { return par[i]; }
  uint32_t &vis_word(unsigned h) { return vw[h]; }
  void hash_clear_vis() { memset(vw, 0xFF, sizeof(vw)); }
};
struct StarRecHostWS
{
  int s[52];
  uint32_t vw[16];
  int &star(int i) { return s[i]; }
  uint32_t &vis_word(unsigned h) { return vw[h]; }
  void hash_clear_vis() { memset(vw, 0xFF, sizeof(vw)); }
};
struct NbrHostWS
{
  int u[36], t[36];
  uint32_t nw[16];
  int &nu(int i) { return u[i]; }
  int &nt(int i) { return t[i]; }
  uint32_t &nbr_word(unsigned h) { return nw[h]; }
  void hash_clear_nbr() { memset(nw, 0xFF, sizeof(nw)); }
};
struct CandVec
{
  std::vector<std::pair<int, int> > v;
  void operator()(int k, int u, int t) { if ((int)v.size() <= k) v.resize(k + 1); v[k] = std::make_pair(u, t); }
};
struct CandRead
{
  const std::vector<std::pair<int, int> > *v;
  void operator()(int k, int &u, int &t) const { u = (*v)[k].first; t = (*v)[k].second; }
};

struct Topo
{
  HostWS big;
  HashWS fast;
  bool use_fast = true;
  int nn = 0;
  float cmin[3], cmax[3];   // bbox from the star's circumcenters (k_cell_bfs), when cc is given
  int run(int site, int t0, const int4 *tets, const float4 *cc = nullptr, const WalkRec *walk = nullptr)
  {
    int ns;
    use_fast = true;
    int st;
    for (int d = 0; d < 3; d++) { cmin[d] = INFINITY; cmax[d] = -INFINITY; }
    if (cc) {
      // k_cell_bfs (star + candidates) followed by k_cell_nbrs (dedup), copied into `fast`
      StarHostWS sw;
      NbrHostWS nw;
      CandVec cands;
      st = star_bfs_cands(site, t0, tets, cc, sw, 52, &ns, cmin, cmax, cands);
      if (walk) {
        // k_cell_bfs runs the walk-record version: same status, same star in the same order, same candidates, same box
        StarRecHostWS rw;
        CandVec rc;
        float rmin[3] = {INFINITY, INFINITY, INFINITY}, rmax[3] = {-INFINITY, -INFINITY, -INFINITY};
        int rns = 0;
        const int rst = star_bfs_rec(site, t0, tets, walk, rw, 52, &rns, rmin, rmax, rc);
        bool same = rst == st;
        if (same && st == CELL_OK) {
          same = rns == ns && rc.v == cands.v && !memcmp(rmin, cmin, 12) && !memcmp(rmax, cmax, 12);
          for (int k = 0; same && k < ns; k++) same = (rw.s[k] & TB_STAR_TET_MASK) == sw.s[k];
        }
        if (!same) { fprintf(stderr, "emul: record BFS differs from vertex BFS at site %d (status %d/%d, ns %d/%d)\n", site, rst, st, rns, ns); abort(); }
      }
      if (st == CELL_OK) {
        CandRead rd{&cands.v};
        nn = nbrs_from_cands(nw, ns + 2, 36, rd);
        if (nn < 0) st = CELL_OVERFLOW;
        else {
          for (int k = 0; k < ns; k++) fast.s[k] = sw.s[k];
          for (int k = 0; k < nn; k++) {
            fast.u[k] = nw.u[k];
            int idx = -1;
            for (int q = 0; q < ns; q++) if (sw.s[q] == nw.t[k]) idx = q;
            fast.nti[k] = (unsigned char)idx;
          }
        }
      }
      // the parent-trick BFS must agree with the plain hashed one
      HashWS chk;
      int ns2, nn2;
      int st2 = star_and_neighbors_hashed(site, t0, tets, chk, 52, 36, &ns2, &nn2);
      // (a walk that overflows its workspace is redone by the general walk whatever the other variant said: an
      // incomplete star can be reported as overflow by the variant that fills a bucket before it meets the hull)
      const bool redo = st == CELL_OVERFLOW || st2 == CELL_OVERFLOW;
      if (!redo && (st2 != st || (st == CELL_OK && (ns2 != ns || nn2 != nn)))) { fprintf(stderr, "emul: BFS variants disagree: st %d/%d ns %d/%d nn %d/%d\n", st, st2, ns, ns2, nn, nn2); abort(); }
      if (st == CELL_OK && st2 == CELL_OK)
        for (int k = 0; k < nn; k++)
          if (chk.nu(k) != fast.nu(k) || chk.nt(k) != fast.nt(k)) { fprintf(stderr, "emul: BFS variants disagree on faces\n"); abort(); }
    } else {
      st = star_and_neighbors_hashed(site, t0, tets, fast, 52, 36, &ns, &nn);
    }
    if (st == CELL_OVERFLOW) {
      use_fast = false;
      st = star_and_neighbors(site, t0, tets, big, 4096, 1024, &ns, &nn);
      if (cc && st == CELL_OK) {
        for (int d = 0; d < 3; d++) { cmin[d] = INFINITY; cmax[d] = -INFINITY; }
        for (int k = 0; k < ns; k++) {
          float4 c = cc[big.star(k)];
          float v[3] = {c.x, c.y, c.z};
          for (int d = 0; d < 3; d++) { cmin[d] = fminf(cmin[d], v[d]); cmax[d] = fmaxf(cmax[d], v[d]); }
        }
      }
    }
    return st;
  }
  int nu(int k) { return use_fast ? fast.nu(k) : big.nu(k); }
  int nt(int k) { return use_fast ? fast.nt(k) : big.nt(k); }
};

struct Rec { uint64_t key, data; };
struct VecEmit
{
  std::vector<Rec> *v;
  void operator()(uint64_t k, uint64_t d) { v->push_back(Rec{k, d}); }
};

// keeps the records of points outside the emitting block (key bit `remote`): RemoteOnlyEmit of kernels.cuh
template <class Inner>
struct RemoteOnly
{
  Inner &in;
  KeyLayout kl;
  void operator()(uint64_t k, uint64_t d)
  {
    if ((k >> (kl.z_bits + kl.cell_bits)) & 1ull) in(k, d);
  }
};

struct BitsIn
{
  const std::vector<unsigned char> *bits;
  int nx, ny;
  bool operator()(int i, int j, int k) const { return (*bits)[((size_t)k * ny + j) * nx + i] != 0; }
};

struct RowIn
{
  const std::vector<unsigned char> *bits;
  int nx, ny;
  uint32_t operator()(int j, int k) const
  {
    uint32_t r = 0;
    for (int i = 0; i < nx; i++) r |= (uint32_t)((*bits)[((size_t)k * ny + j) * nx + i] != 0) << i;
    return r;
  }
};

struct LineEmit
{
  const std::vector<BlockBox> *boxes;
  KeyLayout kl;
  int project, e;
  uint32_t cell;
  const int *lo;
  float value;
  VecEmit *emit; // null: count only
  int nrec;
  void operator()(int yi, int zi, int min_xi, int max_xi)
  {
    std::vector<Rec> tmp;
    VecEmit te{&tmp};
    emit_line(boxes->data(), (int)boxes->size(), e, kl, project, cell, lo[0] + min_xi, lo[0] + max_xi, lo[1] + yi, lo[2] + zi, 0, value, te);
    nrec += (int)tmp.size();
    if (emit) for (auto &r : tmp) (*emit)(r.key, r.data);
  }
};

extern "C" void emu_fill_vert_to_tet(int num_particles, int num_tets, const int *tets, int *v2t)
{
  for (int p = 0; p < num_particles; p++) v2t[p] = -1;
  for (int t = 0; t < num_tets; t++)
    for (int v = 0; v < 4; v++) v2t[tets[8 * t + v]] = std::max(v2t[tets[8 * t + v]], t); // atomicMax in k_vert_to_tet
}

extern "C" void emu_circumcenters(int num_tets, const int *tets, const float *particles, float *out)
{
  for (int t = 0; t < num_tets; t++) {
    const int *v = &tets[8 * t];
    circumcenter(&particles[3 * v[0]], &particles[3 * v[1]], &particles[3 * v[2]], &particles[3 * v[3]], &out[3 * t]);
  }
}

static std::vector<float4> make_cc(int num_tets, const int *tets, const float *particles)
{
  std::vector<float4> cc(num_tets);
  for (int t = 0; t < num_tets; t++) {
    const int *v = &tets[8 * t];
    float o[3], det;
    circumcenter(&particles[3 * v[0]], &particles[3 * v[1]], &particles[3 * v[2]], &particles[3 * v[3]], o, &det);
    cc[t] = float4{o[0], o[1], o[2], fdiv(fabsf(det), 6.0f)};
  }
  return cc;
}

extern "C" void emu_complete(int num_verts, int num_tets, const int *tets, const int *v2t, int *out)
{
  (void)num_tets;
  Topo tp;
  for (int v = 0; v < num_verts; v++) {
    if (v2t[v] < 0) { out[v] = -1; continue; }
    int st = tp.run(v, v2t[v], (const int4 *)tets);
    out[v] = st == CELL_OK ? 1 : 0;
  }
}

extern "C" void emu_volumes(int num_verts, int num_tets, const int *tets, const float *particles, const int *v2t, float *out)
{
  std::vector<float4> cc = make_cc(num_tets, tets, particles);
  Topo tp;
  for (int v = 0; v < num_verts; v++) {
    if (v2t[v] < 0) { out[v] = -2.0f; continue; }
    int st = tp.run(v, v2t[v], (const int4 *)tets);
    if (st != CELL_OK) { out[v] = -1.0f; continue; }
    float vol = 0.0f;
    for (int k = 0; k < tp.nn; k++) {
      AreaAccum aa;
      aa.area = 0.0f;
      int u = tp.nu(k);
      walk_edge_link(v, u, tp.nt(k), (const int4 *)tets, cc.data(), aa);
      float n = 0.0f;
      for (int d = 0; d < 3; d++) {
        float df = fsub(particles[3 * u + d], particles[3 * v + d]);
        n = fadd(n, fmul(df, df));
      }
      vol = fadd(vol, fdiv(fmul(aa.area, fsqrt(n)), 6.0f));
    }
    out[v] = vol;
  }
}

// alg 2 (DTFE): k_vertex_density + k_dtfe_raster single-stepped
struct NoSinkH { void operator()(int, int, int) {} };
static void emu_dtfe_block(const emu_block_t &b, const int *v2t, const BlockBox &bx, const GridGeom &g)
{
  std::vector<float4> cc = make_cc(b.num_tets, b.tets, b.particles);
  std::vector<float> rho(b.num_particles, -1.0f);
  HostWS big;
  for (int v = 0; v < b.num_particles; v++) {
    if (v2t[v] < 0) continue;
    StarHostWS sw;
    NoSinkH sink;
    int ns;
    float cmin[3] = {0, 0, 0}, cmax[3] = {0, 0, 0};
    double sum = 0.0;
    int st = star_bfs_cands(v, v2t[v], (const int4 *)b.tets, cc.data(), sw, 52, &ns, cmin, cmax, sink, &sum);
    if (st == CELL_OVERFLOW) {
      int nn;
      st = star_and_neighbors(v, v2t[v], (const int4 *)b.tets, big, 4096, 1024, &ns, &nn);
      sum = 0.0;
      if (st == CELL_OK) for (int k = 0; k < ns; k++) sum += (double)cc[big.star(k)].w;
    }
    if (st == CELL_OK && sum > 0.0) rho[v] = (float)(4.0 * (double)g.mass / sum);
  }
  for (int t = 0; t < b.num_tets; t++) {
    const int *tv = &b.tets[8 * t];
    float r[4] = {rho[tv[0]], rho[tv[1]], rho[tv[2]], rho[tv[3]]};
    if (r[0] < 0 || r[1] < 0 || r[2] < 0 || r[3] < 0) continue;
    DtfeTet T;
    if (!T.setup(tv, &b.particles[3 * tv[0]], &b.particles[3 * tv[1]], &b.particles[3 * tv[2]], &b.particles[3 * tv[3]], r)) continue;
    int lo[3], hi[3];
    for (int d = 0; d < 3; d++) {
      float mn = b.particles[3 * tv[0] + d], mx = mn;
      for (int j = 1; j < 4; j++) { mn = fminf(mn, b.particles[3 * tv[j] + d]); mx = fmaxf(mx, b.particles[3 * tv[j] + d]); }
      lo[d] = phys2idx1(mn, g.step[d], g.gmin[d]);
      hi[d] = phys2idx1(mx, g.step[d], g.gmin[d]) + 1;
      if (lo[d] < bx.b_lo[d]) lo[d] = bx.b_lo[d];
      if (hi[d] > bx.b_lo[d] + bx.b_num[d] - 1) hi[d] = bx.b_lo[d] + bx.b_num[d] - 1;
    }
    for (int k = lo[2]; k <= hi[2]; k++)
      for (int j = lo[1]; j <= hi[1]; j++)
        for (int i = lo[0]; i <= hi[0]; i++) {
          const float pos[3] = {idx2phys1(i, g.step[0], g.gmin[0]), idx2phys1(j, g.step[1], g.gmin[1]), idx2phys1(k, g.step[2], g.gmin[2])};
          float val;
          if (T.eval(pos, &val)) b.density[((size_t)(k - bx.b_lo[2]) * bx.b_num[1] + (j - bx.b_lo[1])) * bx.b_num[0] + (i - bx.b_lo[0])] = val;
        }
  }
}

extern "C" int emu_dense(emu_params_t *ep, int nblocks, emu_block_t *blocks, int only_gid, const char *outfile, int max_cells)
{
  (void)outfile;
  (void)max_cells;
  tessb200_dense_params p;
  memset(&p, 0, sizeof(p));
  p.alg = ep->alg; p.num_given_bounds = ep->num_given_bounds; p.project = ep->project;
  p.mass = ep->mass; p.eps = ep->eps;
  for (int d = 0; d < 3; d++) {
    p.given_mins[d] = ep->given_mins[d]; p.given_maxs[d] = ep->given_maxs[d];
    p.proj_plane[d] = ep->proj_plane[d]; p.glo_num_idx[d] = ep->glo_num_idx[d];
  }
  for (int i = 0; i < nblocks; i++)
    for (int d = 0; d < 3; d++) {
      if (i == 0 || blocks[i].bounds_min[d] < p.data_mins[d]) p.data_mins[d] = blocks[i].bounds_min[d];
      if (i == 0 || blocks[i].bounds_max[d] > p.data_maxs[d]) p.data_maxs[d] = blocks[i].bounds_max[d];
    }
  grid_step_params(&p);
  GridGeom g;
  for (int d = 0; d < 3; d++) {
    g.gmin[d] = p.grid_phys_mins[d]; g.step[d] = p.grid_step_size[d]; g.dmin[d] = p.data_mins[d]; g.dmax[d] = p.data_maxs[d];
    g.dext_eps[d] = (p.data_maxs[d] - p.data_mins[d]) * 2.0f * FLT_EPSILON; g.gnum[d] = p.glo_num_idx[d];
  }
  g.eps = p.eps; g.mass = p.mass; g.project = p.project ? 1 : 0; g.alg = p.alg;
  g.div = p.project ? g.step[0] * g.step[1] : g.step[0] * g.step[1] * g.step[2];

  std::vector<BlockBox> boxes(nblocks);
  std::vector<uint32_t> cell_base(nblocks);
  long long row_base = 0;
  unsigned long long cb = 0;
  int rc = 0;
  for (int i = 0; i < nblocks; i++) {
    BlockBox &bx = boxes[i];
    block_grid_params(blocks[i].bounds_min, blocks[i].bounds_max, &p, bx.b_lo, bx.b_num);
    phys_box(blocks[i].bounds_min, blocks[i].bounds_max, &p, bx.p_lo, bx.p_hi);
    bx.row_base = row_base;
    long long nrows = p.project ? bx.b_num[1] : (long long)bx.b_num[1] * bx.b_num[2];
    row_base += nrows;
    cell_base[i] = (uint32_t)cb;
    cb += (unsigned long long)blocks[i].num_orig_particles;
    long long npts = nrows * bx.b_num[0];
    for (int d = 0; d < 3; d++) { blocks[i].block_min_idx[d] = bx.b_lo[d]; blocks[i].block_num_idx[d] = bx.b_num[d]; }
    blocks[i].num_grid_pts = (int)npts;
    if (bx.b_num[0] < 1 || bx.b_num[1] < 1 || bx.b_num[2] < 1) rc = -3;   // as api.cu: "block owns no grid points along axis"
    else if (npts > blocks[i].density_capacity) rc = -1;
  }
  if (rc) return rc;
  KeyLayout kl;
  kl.cell_bits = ceil_log2(cb + 1);
  key_z_range(boxes.data(), boxes.size(), p.project, &kl);

  if (p.alg == TESSB200_DENSE_DTFE) {
    for (int bi = 0; bi < nblocks; bi++) {
      emu_block_t &b = blocks[bi];
      memset(b.density, 0, sizeof(float) * (size_t)b.num_grid_pts);
      std::vector<int> v2t_own;
      const int *v2t = b.vert_to_tet;
      if (!v2t) {
        v2t_own.resize(b.num_particles);
        emu_fill_vert_to_tet(b.num_particles, b.num_tets, b.tets, v2t_own.data());
        v2t = v2t_own.data();
      }
      emu_dtfe_block(b, v2t, boxes[bi], g);
    }
    for (int d = 0; d < 3; d++) {
      ep->data_mins[d] = p.data_mins[d]; ep->data_maxs[d] = p.data_maxs[d];
      ep->grid_phys_mins[d] = p.grid_phys_mins[d]; ep->grid_phys_maxs[d] = p.grid_phys_maxs[d];
      ep->grid_step_size[d] = p.grid_step_size[d];
    }
    return 0;
  }
  std::vector<Rec> recs;
  VecEmit emit{&recs};
  std::vector<char> gathered(nblocks, 0);      // blocks whose own CIC deposits are already in their density (the gather)
  const bool cic_gather = getenv("TESSB200_EMUL_CIC_RECORDS") == nullptr;
  Topo tp;
  for (int bi = 0; bi < nblocks; bi++) {
    emu_block_t &b = blocks[bi];
    if (only_gid >= 0 && b.gid != only_gid) continue;
    std::vector<int> v2t_own;
    const int *v2t = b.vert_to_tet;
    if (!v2t) {
      v2t_own.resize(b.num_particles);
      emu_fill_vert_to_tet(b.num_particles, b.num_tets, b.tets, v2t_own.data());
      v2t = v2t_own.data();
    }
    if (p.alg == TESSB200_DENSE_CIC && !g.project && cic_gather) {
      // k_cic_prepare + stable sort by base cell + scan + k_cic_permute + k_cic_gather, in the kernels' terms: the block's own
      // deposits are gathered per grid point (written into b.density here), only the records that leave the block are kept
      const BlockBox &bx = boxes[bi];
      const int o[3] = {bx.b_lo[0] - 1, bx.b_lo[1] - 1, bx.b_lo[2] - 1}, dd[3] = {bx.b_num[0] + 1, bx.b_num[1] + 1, bx.b_num[2] + 1};
      const size_t ncell = (size_t)dd[0] * dd[1] * dd[2];
      const int np = b.num_orig_particles;
      std::vector<float> q8(8 * (size_t)np);
      std::vector<uint32_t> key(np), ids(np);
      std::vector<unsigned int> count(ncell + 1, 0u), start(ncell + 2, 0u);
      RemoteOnly<VecEmit> remote{emit, kl};
      for (int cell = 0; cell < np; cell++) {
        int i0[3];
        float vals[8];
        cic_weights(&b.particles[3 * cell], g.mass, g, i0, vals);
        for (int n = 0; n < 8; n++) q8[8 * (size_t)cell + n] = fdiv(vals[n], g.div);
        const long long cx = (long long)i0[0] - o[0], cy = (long long)i0[1] - o[1], cz = (long long)i0[2] - o[2];
        key[cell] = 0xffffffffu;
        if (cx >= 0 && cx < dd[0] && cy >= 0 && cy < dd[1] && cz >= 0 && cz < dd[2]) {
          key[cell] = (uint32_t)((cz * dd[1] + cy) * dd[0] + cx);
          count[key[cell]]++;
        }
        ids[cell] = (uint32_t)cell;
        int n = 0;
        for (int dz = 0; dz < 2; dz++)
          for (int dy = 0; dy < 2; dy++)
            for (int dx = 0; dx < 2; dx++, n++)
              emit_line(boxes.data(), nblocks, bi, kl, g.project, cell_base[bi] + cell, i0[0] + dx, i0[0] + dx, i0[1] + dy, i0[2] + dz, 1, vals[n], remote);
      }
      std::stable_sort(ids.begin(), ids.end(), [&](uint32_t a, uint32_t c) { return key[a] < key[c]; });
      for (size_t c = 0; c <= ncell; c++) start[c + 1] = start[c] + count[c];          // start[c] = exclusive scan, ncell + 1 entries used
      std::vector<float> q8s(8 * (size_t)np);
      for (int q = 0; q < np; q++) memcpy(&q8s[8 * (size_t)q], &q8[8 * (size_t)ids[q]], 32);
      if (np == 0) { ids.push_back(0); q8s.resize(8); }
      for (int lz = 0; lz < bx.b_num[2]; lz++)
        for (int ly = 0; ly < bx.b_num[1]; ly++)
          for (int lx = 0; lx < bx.b_num[0]; lx++) {
            const int x = bx.b_lo[0] + lx, y = bx.b_lo[1] + ly, z = bx.b_lo[2] + lz;
            float cur = 0.0f;
            if (x >= bx.p_lo[0] && x <= bx.p_hi[0] && y >= bx.p_lo[1] && y <= bx.p_hi[1] && z >= bx.p_lo[2] && z <= bx.p_hi[2]) {
              const long long row = dd[0], slab = (long long)dd[0] * dd[1];
              const unsigned int *cs11 = start.data() + ((long long)lz * slab + (long long)ly * row + lx);
              CicLists l;
              unsigned int total = 0;
              for (int h = 0; h < 4; h++) {
                const int dy = h & 1, dz = h >> 1;
                const unsigned int *cs = cs11 + (dy ? 0 : row) + (dz ? 0 : slab);
                const unsigned int s0 = cs[0], s1 = cs[1], s2 = cs[2];
                const int n1 = h * 2 + 1, n0 = h * 2;
                l.set(n1, s0, s1);
                l.set(n0, s1, s2);
                l.head[n1] = s0 < s1 ? ids[s0] : 0xffffffffu;
                l.head[n0] = s1 < s2 ? ids[s1] : 0xffffffffu;
                total += s2 - s0;
              }
              if (total) cur = cic_merge_sum(l, total, ids.data(), q8s.data());
            }
            b.density[((size_t)lz * bx.b_num[1] + ly) * bx.b_num[0] + lx] = cur;
          }
      gathered[bi] = 1;
      continue;
    }
    if (p.alg == TESSB200_DENSE_CIC) {
      for (int cell = 0; cell < b.num_orig_particles; cell++) {
        int i0[3];
        float vals[8];
        cic_weights(&b.particles[3 * cell], g.mass, g, i0, vals);
        int n = 0;
        for (int dz = 0; dz < 2; dz++)
          for (int dy = 0; dy < 2; dy++)
            for (int dx = 0; dx < 2; dx++, n++)
              emit_line(boxes.data(), nblocks, bi, kl, g.project, cell_base[bi] + cell, i0[0] + dx, i0[0] + dx, i0[1] + dy, i0[2] + dz, 1, vals[n], emit);
      }
      continue;
    }
    std::vector<float4> cc = make_cc(b.num_tets, b.tets, b.particles);
    std::vector<WalkRec> walk(b.num_tets);
    for (int t = 0; t < b.num_tets; t++) {
      WalkRec r;
      for (int q = 0; q < 4; q++) r.nb[q] = b.tets[8 * t + 4 + q];
      r.cx = cc[t].x; r.cy = cc[t].y; r.cz = cc[t].z;
      r.perm = walk_perm(&b.tets[8 * t], &b.tets[8 * t + 4], (const int4 *)b.tets);
      walk[t] = r;
    }
    for (int cell = 0; cell < b.num_orig_particles; cell++) {
      if (v2t[cell] < 0) continue;
      int st = tp.run(cell, v2t[cell], (const int4 *)b.tets, cc.data(), walk.data());
      if (st != CELL_OK) continue;
      const int nn = tp.nn;
      float cmin[3] = {INFINITY, INFINITY, INFINITY}, cmax[3] = {-INFINITY, -INFINITY, -INFINITY};
      const float *site = &b.particles[3 * cell];
      std::vector<float> planes(6 * nn);
      bool bad = false;
      for (int k = 0; k < nn; k++) {
        FaceAccum fa;
        fa.cmin = cmin; fa.cmax = cmax;
        int n = walk_edge_link(cell, tp.nu(k), tp.nt(k), (const int4 *)b.tets, cc.data(), fa);
        if (n < 0) { bad = true; break; }
        {
          // k_cell_faces' walk on the 32-byte records (slot permutations) must visit the same tets
          FaceAccum fb;
          fb.cmin = nullptr; fb.cmax = nullptr;
          const int *v0 = &b.tets[8 * tp.nt(k)];
          int s_c = -1, s_u = -1;
          for (int q = 0; q < 4; q++) { if (v0[q] == cell) s_c = q; if (v0[q] == tp.nu(k)) s_u = q; }
          int n2 = walk_edge_link_rec(s_c, s_u, tp.nt(k), walk.data(), fb);
          if (n2 != n || memcmp(fa.nrm, fb.nrm, 12) || memcmp(fa.v0, fb.v0, 12) || memcmp(fa.prev, fb.prev, 12)) {
            fprintf(stderr, "emul: record walk differs from vertex walk (%d vs %d steps)\n", n2, n);
            abort();
          }
        }
        newell_term(fa.nrm, fa.prev, fa.v0);
        newell_finish(fa.nrm, fa.v0, site);
        for (int d = 0; d < 3; d++) { planes[6 * k + d] = fa.nrm[d]; planes[6 * k + 3 + d] = fa.v0[d]; }
      }
      if (bad) continue;
      for (int d = 0; d < 3; d++)
        if (cmin[d] != tp.cmin[d] || cmax[d] != tp.cmax[d]) { fprintf(stderr, "emul: star bbox != face bbox\n"); abort(); }
      bool outside = false;
      for (int d = 0; d < 3; d++)
        if (cmin[d] < fsub(g.dmin[d], g.dext_eps[d]) || cmax[d] > fadd(g.dmax[d], g.dext_eps[d])) outside = true;
      if (outside) continue;
      int lo[3], n3[3];
      for (int d = 0; d < 3; d++) {
        lo[d] = phys2idx1(cmin[d], g.step[d], g.gmin[d]);
        n3[d] = phys2idx1(cmax[d], g.step[d], g.gmin[d]) - lo[d] + 1;
      }
      std::vector<unsigned char> bits((size_t)n3[0] * n3[1] * n3[2]);
      for (int k = 0; k < n3[2]; k++)
        for (int j = 0; j < n3[1]; j++)
          for (int i = 0; i < n3[0]; i++) {
            float pt[3] = {fadd(idx2phys1(lo[0], g.step[0], g.gmin[0]), fmul((float)i, g.step[0])),
                           fadd(idx2phys1(lo[1], g.step[1], g.gmin[1]), fmul((float)j, g.step[1])),
                           fadd(idx2phys1(lo[2], g.step[2], g.gmin[2]), fmul((float)k, g.step[2]))};
            bool pos = false, neg = false;
            for (int f = 0; f < nn && !(pos && neg); f++) {
              int sd = plane_side(&planes[6 * f], &planes[6 * f + 3], pt, g.eps);
              pos |= sd > 0;
              neg |= sd < 0;
            }
            bits[((size_t)k * n3[1] + j) * n3[0] + i] = !(pos && neg);
          }
      BitsIn inside{&bits, n3[0], n3[1]};
      RowIn row{&bits, n3[0], n3[1]};
      const bool bitpath = n3[0] <= 32;   // the kernels' bit-parallel walk (scan_cell_bits), else the generic one
      LineEmit cnt{&boxes, kl, g.project, bi, cell_base[bi] + (uint32_t)cell, lo, 0.0f, nullptr, 0};
      int tot = bitpath ? scan_cell_bits(n3[0], n3[1], n3[2], row, cnt) : scan_cell(n3[0], n3[1], n3[2], inside, cnt);
      if (tot > 0) {
        LineEmit le{&boxes, kl, g.project, bi, cell_base[bi] + (uint32_t)cell, lo, fdiv(g.mass, (float)tot), &emit, 0};
        if (bitpath) scan_cell_bits(n3[0], n3[1], n3[2], row, le);
        else scan_cell(n3[0], n3[1], n3[2], inside, le);
      } else {
        int i0[3];
        float vals[8];
        cic_weights(site, g.mass, g, i0, vals);
        int n = 0;
        for (int dz = 0; dz < 2; dz++)
          for (int dy = 0; dy < 2; dy++)
            for (int dx = 0; dx < 2; dx++, n++)
              emit_line(boxes.data(), nblocks, bi, kl, g.project, cell_base[bi] + cell, i0[0] + dx, i0[0] + dx, i0[1] + dy, i0[2] + dz, 0, vals[n], emit);
      }
    }
  }
  std::sort(recs.begin(), recs.end(), [](const Rec &a, const Rec &b) { return a.key < b.key; });
  for (int i = 0; i < nblocks; i++)
    if (!gathered[i]) memset(blocks[i].density, 0, sizeof(float) * (size_t)blocks[i].num_grid_pts);
  for (const Rec &r : recs) {
    long long row = (long long)key_row(kl, r.key);
    int bi = nblocks - 1;
    while (bi > 0 && boxes[bi].row_base > row) bi--;
    int nx = boxes[bi].b_num[0];
    float *dst = blocks[bi].density + (row - boxes[bi].row_base) * nx;
    int x0 = (int)(r.data & 0xffffu), len = (int)((r.data >> 16) & 0x7fffu), fp = (int)((r.data >> 31) & 1u);
    float m = u2f((uint32_t)(r.data >> 32));
    for (int x = 0; x < len; x++) dst[x0 + x] = accumulate(dst[x0 + x], m, g.div, fp);
  }
  for (int d = 0; d < 3; d++) {
    ep->data_mins[d] = p.data_mins[d]; ep->data_maxs[d] = p.data_maxs[d];
    ep->grid_phys_mins[d] = p.grid_phys_mins[d]; ep->grid_phys_maxs[d] = p.grid_phys_maxs[d];
    ep->grid_step_size[d] = p.grid_step_size[d];
  }
  ep->seconds = 0;
  return 0;
}

// k_cic_gather's sum at one grid point: the eight lists [pos[n], end[n]) of sorted_ids, weights vals[8 * id + n]
extern "C" int emu_cic_point(const uint32_t *sorted_ids, const float *vals, const unsigned int *pos, const unsigned int *end, float *out)
{
  CicLists l;
  unsigned int total = 0;
  for (int n = 0; n < 8; n++) {
    l.set(n, pos[n], end[n]);
    total += end[n] - pos[n];
    l.head[n] = pos[n] < end[n] ? sorted_ids[pos[n]] : 0xffffffffu;
  }
  out[0] = cic_merge_sum(l, total, sorted_ids, vals);
  // one step past the end must be the neutral element
  out[1] = fadd(out[0], cic_merge_step(l, sorted_ids, vals));
  return 0;
}
