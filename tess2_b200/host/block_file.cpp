// block_file.cpp -- tess2's hand-off file between the two stages ("del.out"), SURVEY 8(f) N3.
//
// tess_save (src/tess.cpp:126-137) writes the blocks with diy::io::write_blocks and the callback
// save_block_light (src/tess.cpp:198-221); examples/dense/main.cpp:158-161 reads them back with
// diy::io::read_blocks + load_block_light (src/tess.cpp:223-259).  DIY is not vendored under the
// reference and has no pinned revision (README.md:22), so the container format below restates DIY's
// published block-file layout (diy/io/block.hpp) -- PARITY UNPINNED -- while the block payload is
// pinned by the reference's own field order:
//
//   file    := block buffer * nblocks , footer
//   footer  := vector<GidOffsetCount>{ size_t n ; n x { int gid ; (pad 4) ; int64 offset ; int64 count } }   (sorted by gid)
//              , extra (opaque bytes, diy::MemoryBuffer) , size_t footer_size        (the last 8 bytes of the file)
//   block   := link record (DIY version dependent, skipped on read) , payload
//   payload := int gid ; Bounds bounds ; Bounds box ; Bounds data_bounds ; int num_orig_particles ; int num_particles ;
//              float particles[3 np] ; int rem_gids[np - norig] ; int rem_lids[np - norig] ; int num_grid_pts ;
//              float density[num_grid_pts] ; int complete ; int num_tets ; tet_t tets[num_tets] ; int vert_to_tet[np]
//   Bounds  := DYNAMIC: size_t 3, float min[3], size_t 3, float max[3]  (DIY with DynamicPoint: what
//              `diy::ContinuousBounds bounds { 3 }`, include/tess/delaunay.hpp:19-21, compiles against)
//            | STATIC4: float min[4], float max[4]                       (DIY with DIY_MAX_DIM = 4 points, raw struct copy)
//
// The reader does not interpret the link record: it finds the payload as the first position whose
// int equals the footer's gid and from which the payload parses to EXACTLY the end of the block
// buffer with consistent counts, trying both Bounds layouts.  That makes it independent of the link
// class and of DIY's link serialisation.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>

#include "../../include/tess_b200_host.h"

void tessb200_host_set_error(const std::string &s);   // host_tess.cpp

namespace
{

struct GidOffsetCount { int gid; int pad; int64_t offset; int64_t count; };
static_assert(sizeof(GidOffsetCount) == 24, "DIY's GidOffsetCount is {int, offset_t, offset_t} with natural alignment");

struct Writer {
  std::vector<char> buf;
  template <class T> void put(const T &v) { const char *p = (const char *)&v; buf.insert(buf.end(), p, p + sizeof(T)); }
  void raw(const void *p, size_t n) { if (n) buf.insert(buf.end(), (const char *)p, (const char *)p + n); }
  void bounds(const float *mn, const float *mx, int layout)
  {
    if (layout == TESSB200_DIY_BOUNDS_DYNAMIC) {
      put<uint64_t>(3); raw(mn, 12);
      put<uint64_t>(3); raw(mx, 12);
    } else {
      const float z = 0.0f;
      raw(mn, 12); put(z);
      raw(mx, 12); put(z);
    }
  }
};

struct Reader {
  const char *p, *end;
  bool ok = true;
  template <class T> T get()
  {
    T v{};
    if ((size_t)(end - p) < sizeof(T)) { ok = false; return v; }
    memcpy(&v, p, sizeof(T));
    p += sizeof(T);
    return v;
  }
  const char *span(size_t n)
  {
    if (!ok || (size_t)(end - p) < n) { ok = false; return nullptr; }
    const char *q = p;
    p += n;
    return q;
  }
  void bounds(float *mn, float *mx, int layout)
  {
    if (layout == TESSB200_DIY_BOUNDS_DYNAMIC) {
      if (get<uint64_t>() != 3) ok = false;
      const char *a = span(12);
      if (get<uint64_t>() != 3) ok = false;
      const char *b = span(12);
      if (ok) { memcpy(mn, a, 12); memcpy(mx, b, 12); }
    } else {
      const char *a = span(16), *b = span(16);
      if (ok) { memcpy(mn, a, 12); memcpy(mx, b, 12); }
    }
  }
};

// payload at [p, end) in the given Bounds layout; true iff it parses to exactly `end`
bool parse_payload(const char *p, const char *end, int gid, int layout, tessb200_host_dblock *out, bool fill)
{
  Reader r{p, end};
  tessb200_host_dblock b;
  memset(&b, 0, sizeof(b));
  b.gid = r.get<int>();
  if (!r.ok || b.gid != gid) return false;
  r.bounds(b.bounds_min, b.bounds_max, layout);
  r.bounds(b.box_min, b.box_max, layout);
  r.bounds(b.data_min, b.data_max, layout);
  b.num_orig_particles = r.get<int>();
  b.num_particles = r.get<int>();
  if (!r.ok || b.num_orig_particles < 0 || b.num_particles < b.num_orig_particles) return false;
  const size_t np = (size_t)b.num_particles, ng = np - (size_t)b.num_orig_particles;
  const char *particles = r.span(12 * np), *rem_gids = r.span(4 * ng), *rem_lids = r.span(4 * ng);
  b.num_grid_pts = r.get<int>();
  if (!r.ok || b.num_grid_pts < 0) return false;
  const char *density = r.span(4 * (size_t)b.num_grid_pts);
  b.complete = r.get<int>();
  b.num_tets = r.get<int>();
  if (!r.ok || b.num_tets < 0) return false;
  const char *tets = r.span(32 * (size_t)b.num_tets), *v2t = r.span(4 * np);
  if (!r.ok || r.p != end) return false;
  if (!fill) return true;
  // malloc as load_block_light does (density is new[] there; a C ABI cannot hand that out)
  auto dup = [](const char *src, size_t n) -> void * {
    void *q = malloc(n ? n : 1);
    if (q && n) memcpy(q, src, n);
    return q;
  };
  b.particles = (float *)dup(particles, 12 * np);
  b.rem_gids = (int *)dup(rem_gids, 4 * ng);
  b.rem_lids = (int *)dup(rem_lids, 4 * ng);
  b.density = (float *)dup(density, 4 * (size_t)b.num_grid_pts);
  b.tets = (int *)dup(tets, 32 * (size_t)b.num_tets);
  b.vert_to_tet = (int *)dup(v2t, 4 * np);
  if (!b.particles || !b.rem_gids || !b.rem_lids || !b.density || !b.tets || !b.vert_to_tet) {
    free(b.particles); free(b.rem_gids); free(b.rem_lids); free(b.density); free(b.tets); free(b.vert_to_tet);
    return false;
  }
  *out = b;
  return true;
}

bool closed_boxes_touch(const tessb200_host_dblock &a, const tessb200_host_dblock &b)
{
  for (int d = 0; d < 3; d++)
    if (a.bounds_max[d] < b.bounds_min[d] || b.bounds_max[d] < a.bounds_min[d]) return false;
  return true;
}

struct File {
  FILE *f = nullptr;
  ~File() { if (f) fclose(f); }
};

}  // namespace

extern "C" void tessb200_host_free_dblocks(int nblocks, tessb200_host_dblock *blocks)
{
  if (!blocks) return;
  for (int i = 0; i < nblocks; i++) {
    free(blocks[i].particles); free(blocks[i].rem_gids); free(blocks[i].rem_lids);
    free(blocks[i].density); free(blocks[i].tets); free(blocks[i].vert_to_tet);
  }
  free(blocks);
}

static int write_blocks(const char *path, int nblocks, const tessb200_host_dblock *blocks, int bounds_layout, const void *extra, size_t extra_size)
{
  if (!path || nblocks < 0 || (nblocks && !blocks) || (extra_size && !extra) ||
      (bounds_layout != TESSB200_DIY_BOUNDS_DYNAMIC && bounds_layout != TESSB200_DIY_BOUNDS_STATIC4)) {
    tessb200_host_set_error("write_blocks: bad argument");
    return -1;
  }
  for (int i = 0; i < nblocks; i++) {
    const tessb200_host_dblock &b = blocks[i];
    const bool bad = b.num_orig_particles < 0 || b.num_particles < b.num_orig_particles || b.num_tets < 0 || b.num_grid_pts < 0 ||
                     (b.num_particles && (!b.particles || !b.vert_to_tet)) || (b.num_tets && !b.tets) || (b.num_grid_pts && !b.density);
    if (bad) { tessb200_host_set_error("write_blocks: inconsistent block " + std::to_string(i)); return -1; }
  }
  File fh;
  fh.f = fopen(path, "wb");
  if (!fh.f) { tessb200_host_set_error(std::string("write_blocks: cannot open ") + path); return -4; }
  std::vector<GidOffsetCount> toc;
  int64_t offset = 0;
  for (int i = 0; i < nblocks; i++) {
    const tessb200_host_dblock &b = blocks[i];
    Writer w;
    // link record: the base diy::Link (type id string, then the neighbour list {gid, proc}); every
    // block whose closed bounds touch this one's is a neighbour, all on process 0
    const std::string id = "N3diy4LinkE";
    w.put<uint64_t>(id.size());
    w.raw(id.data(), id.size());
    std::vector<int> nbrs;
    for (int j = 0; j < nblocks; j++)
      if (j != i && closed_boxes_touch(b, blocks[j])) { nbrs.push_back(blocks[j].gid); nbrs.push_back(0); }
    w.put<uint64_t>(nbrs.size() / 2);
    w.raw(nbrs.data(), nbrs.size() * sizeof(int));
    // payload, save_block_light's order (src/tess.cpp:198-221)
    const size_t np = (size_t)b.num_particles, ng = np - (size_t)b.num_orig_particles;
    w.put(b.gid);
    w.bounds(b.bounds_min, b.bounds_max, bounds_layout);
    w.bounds(b.box_min, b.box_max, bounds_layout);
    w.bounds(b.data_min, b.data_max, bounds_layout);
    w.put(b.num_orig_particles);
    w.put(b.num_particles);
    if (fwrite(w.buf.data(), 1, w.buf.size(), fh.f) != w.buf.size()) goto io_error;
    {
      int64_t count = (int64_t)w.buf.size();
      auto out = [&](const void *p, size_t n) { count += (int64_t)n; return n == 0 || fwrite(p, 1, n, fh.f) == n; };
      std::vector<int> zeros;
      const int *rg = b.rem_gids, *rl = b.rem_lids;
      if (ng && (!rg || !rl)) { zeros.assign(ng, -1); if (!rg) rg = zeros.data(); if (!rl) rl = zeros.data(); }
      if (!out(b.particles, 12 * np) || !out(rg, 4 * ng) || !out(rl, 4 * ng) || !out(&b.num_grid_pts, 4) ||
          !out(b.density, 4 * (size_t)b.num_grid_pts) || !out(&b.complete, 4) || !out(&b.num_tets, 4) ||
          !out(b.tets, 32 * (size_t)b.num_tets) || !out(b.vert_to_tet, 4 * np))
        goto io_error;
      toc.push_back(GidOffsetCount{b.gid, 0, offset, count});
      offset += count;
    }
  }
  {
    std::sort(toc.begin(), toc.end(), [](const GidOffsetCount &a, const GidOffsetCount &b) { return a.gid < b.gid; });
    Writer w;
    w.put<uint64_t>(toc.size());
    w.raw(toc.data(), toc.size() * sizeof(GidOffsetCount));
    w.put<uint64_t>(0);             // extra: diy::MemoryBuffer = read position, then the bytes as vector<char>
    w.put<uint64_t>(extra_size);
    w.raw(extra, extra_size);
    const uint64_t footer_size = w.buf.size();
    w.put(footer_size);
    if (fwrite(w.buf.data(), 1, w.buf.size(), fh.f) != w.buf.size()) goto io_error;
  }
  if (fflush(fh.f)) goto io_error;
  return 0;
io_error:
  tessb200_host_set_error(std::string("write_blocks: write to ") + path + " failed");
  return -4;
}

static int read_blocks(const char *path, int *nblocks, tessb200_host_dblock **blocks, int *bounds_layout)
{
  if (!path || !nblocks || !blocks) { tessb200_host_set_error("read_blocks: NULL argument"); return -1; }
  *nblocks = 0;
  *blocks = nullptr;
  File fh;
  fh.f = fopen(path, "rb");
  if (!fh.f) { tessb200_host_set_error(std::string("read_blocks: cannot open ") + path); return -4; }
  if (fseeko(fh.f, 0, SEEK_END)) { tessb200_host_set_error("read_blocks: seek failed"); return -4; }
  const int64_t fsize = (int64_t)ftello(fh.f);
  auto read_at = [&](int64_t off, void *dst, size_t n) { return fseeko(fh.f, (off_t)off, SEEK_SET) == 0 && fread(dst, 1, n, fh.f) == n; };
  uint64_t footer_size = 0;
  if (fsize < 16 || !read_at(fsize - 8, &footer_size, 8) || footer_size < 8 || footer_size > (uint64_t)(fsize - 8)) {
    tessb200_host_set_error("read_blocks: not a DIY block file (no footer)");
    return -5;
  }
  const int64_t footer_offset = fsize - 8 - (int64_t)footer_size;
  std::vector<char> footer(footer_size);
  if (!read_at(footer_offset, footer.data(), footer.size())) { tessb200_host_set_error("read_blocks: short read (footer)"); return -4; }
  uint64_t n = 0;
  memcpy(&n, footer.data(), 8);
  if (n > (footer_size - 8) / sizeof(GidOffsetCount) || n > 0x7fffffffu) { tessb200_host_set_error("read_blocks: corrupt footer (block count)"); return -5; }
  std::vector<GidOffsetCount> toc(n);
  if (n) memcpy(toc.data(), footer.data() + 8, n * sizeof(GidOffsetCount));
  for (const GidOffsetCount &e : toc)
    if (e.offset < 0 || e.count < 4 || e.offset > footer_offset || e.count > footer_offset - e.offset) {
      tessb200_host_set_error("read_blocks: corrupt footer (block extent)");
      return -5;
    }
  tessb200_host_dblock *out = (tessb200_host_dblock *)calloc(n ? n : 1, sizeof(tessb200_host_dblock));
  if (!out) { tessb200_host_set_error("out of memory"); return -2; }
  int layout_seen = -1;
  std::vector<char> buf;
  for (size_t i = 0; i < n; i++) {
    const GidOffsetCount &e = toc[i];
    try { buf.resize((size_t)e.count); } catch (const std::exception &) { tessb200_host_free_dblocks((int)i, out); tessb200_host_set_error("out of memory"); return -2; }
    if (!read_at(e.offset, buf.data(), buf.size())) { tessb200_host_free_dblocks((int)i, out); tessb200_host_set_error("read_blocks: short read (block)"); return -4; }
    const char *p = buf.data(), *end = p + buf.size();
    // the link record is small (tens of bytes per neighbour): the payload starts within the first 64 KiB
    const size_t scan = std::min<size_t>(buf.size() - 4, 65536);
    bool found = false;
    const int first = layout_seen >= 0 ? layout_seen : TESSB200_DIY_BOUNDS_DYNAMIC;
    for (int t = 0; t < 2 && !found; t++) {
      const int layout = t == 0 ? first : 1 - first;
      for (size_t s = 0; s <= scan && !found; s++) {
        int g;
        memcpy(&g, p + s, 4);
        if (g != e.gid || !parse_payload(p + s, end, e.gid, layout, nullptr, false)) continue;
        if (!parse_payload(p + s, end, e.gid, layout, &out[i], true)) { tessb200_host_free_dblocks((int)i, out); tessb200_host_set_error("out of memory"); return -2; }
        found = true;
        layout_seen = layout;
      }
    }
    if (!found) {
      tessb200_host_free_dblocks((int)i, out);
      tessb200_host_set_error("read_blocks: block gid " + std::to_string(e.gid) + " does not hold a save_block_light payload");
      return -5;
    }
  }
  *nblocks = (int)n;
  *blocks = out;
  if (bounds_layout) *bounds_layout = layout_seen < 0 ? TESSB200_DIY_BOUNDS_DYNAMIC : layout_seen;
  return 0;
}

// no exception crosses the C ABI
extern "C" int tessb200_host_write_blocks(const char *path, int nblocks, const tessb200_host_dblock *blocks, int bounds_layout,
                                          const void *extra, size_t extra_size)
{
  try {
    return write_blocks(path, nblocks, blocks, bounds_layout, extra, extra_size);
  } catch (const std::exception &e) {
    tessb200_host_set_error(std::string("write_blocks: ") + e.what());
    return -2;
  }
}

extern "C" int tessb200_host_read_blocks(const char *path, int *nblocks, tessb200_host_dblock **blocks, int *bounds_layout)
{
  try {
    return read_blocks(path, nblocks, blocks, bounds_layout);
  } catch (const std::exception &e) {
    tessb200_host_set_error(std::string("read_blocks: ") + e.what());
    if (nblocks) *nblocks = 0;
    return -2;
  }
}
