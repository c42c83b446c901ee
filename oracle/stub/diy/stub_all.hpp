// Single-process stand-in for the DIY block-parallel library, TEST INFRASTRUCTURE ONLY.
//
// DIY (github.com/diatomic/diy) is a header-only dependency of tess2 that is not
// vendored under /root/reference and has no pinned revision there (README.md:22
// clones HEAD).  This stub provides exactly the surface that the reference's
// include/tess/{tess,dense,delaunay,tet}.hpp and src/{dense,tet,volume}.cpp touch,
// so that those files compile UNMODIFIED into oracle/_ref/.  Semantics kept:
//   * Master::foreach visits local blocks in insertion order,
//   * Proxy::enqueue appends raw bytes to a per-destination-gid queue,
//   * Master::exchange moves every queue to the destination block,
//   * Proxy::incoming(v) lists source gids in ascending order (std::map order,
//     as DIY's IncomingQueues map does),
//   * diy::in(link, p, out, domain) = closed-box containment of p in each
//     neighbour's (wrap-adjusted) bounds, DIY pick.hpp's distance()==0 test.
// Nothing in the product (tess2_b200/) includes this file.
#ifndef TESSB200_ORACLE_DIY_STUB_HPP
#define TESSB200_ORACLE_DIY_STUB_HPP

#include <vector>
#include <map>
#include <string>
#include <cstring>
#include <cstdlib>
#include <cstdio>
#include <functional>
#include "mpi.h"

namespace diy
{
  // ---- geometry types ----
  template <class C, unsigned D>
  struct Point
  {
    C x[D];
    Point() { for (unsigned i = 0; i < D; i++) x[i] = 0; }
    Point(const C *a) { for (unsigned i = 0; i < D; i++) x[i] = a[i]; }
    C &operator[](unsigned i) { return x[i]; }
    const C &operator[](unsigned i) const { return x[i]; }
  };

  struct Direction
  {
    int x[4];
    Direction() { x[0] = x[1] = x[2] = x[3] = 0; }
    int &operator[](int i) { return x[i]; }
    const int &operator[](int i) const { return x[i]; }
    bool operator==(const Direction &o) const { return x[0] == o.x[0] && x[1] == o.x[1] && x[2] == o.x[2]; }
  };

  struct ContinuousBounds
  {
    float min[4], max[4];
    ContinuousBounds() { for (int i = 0; i < 4; i++) min[i] = max[i] = 0; }
    ContinuousBounds(int) { for (int i = 0; i < 4; i++) min[i] = max[i] = 0; }
  };
  struct DiscreteBounds
  {
    int min[4], max[4];
    DiscreteBounds() { for (int i = 0; i < 4; i++) min[i] = max[i] = 0; }
    DiscreteBounds(int) { for (int i = 0; i < 4; i++) min[i] = max[i] = 0; }
  };

  struct BlockID { int gid, proc; };

  // ---- serialization ----
  struct BinaryBuffer
  {
    virtual ~BinaryBuffer() {}
    virtual void save_binary(const char *x, size_t count) = 0;
    virtual void load_binary(char *x, size_t count) = 0;
  };
  struct MemoryBuffer : public BinaryBuffer
  {
    std::vector<char> buffer;
    size_t position;
    MemoryBuffer() : position(0) {}
    void save_binary(const char *x, size_t count) override
    {
      buffer.insert(buffer.end(), x, x + count);
      position = buffer.size();
    }
    void load_binary(char *x, size_t count) override
    {
      memcpy(x, &buffer[position], count);
      position += count;
    }
    size_t size() const { return buffer.size(); }
    void reset() { position = 0; }
    void clear() { buffer.clear(); position = 0; }
  };

  template <class T>
  struct Serialization
  {
    static void save(BinaryBuffer &bb, const T &x) { bb.save_binary((const char *)&x, sizeof(T)); }
    static void load(BinaryBuffer &bb, T &x) { bb.load_binary((char *)&x, sizeof(T)); }
  };
  template <class T> void save(BinaryBuffer &bb, const T &x) { Serialization<T>::save(bb, x); }
  template <class T> void load(BinaryBuffer &bb, T &x) { Serialization<T>::load(bb, x); }
  template <class T> void save(BinaryBuffer &bb, const T *x, size_t n) { if (n) bb.save_binary((const char *)x, sizeof(T) * n); }
  template <class T> void load(BinaryBuffer &bb, T *x, size_t n) { if (n) bb.load_binary((char *)x, sizeof(T) * n); }

  // ---- links ----
  struct Link
  {
    std::vector<BlockID> neighbors_;
    virtual ~Link() {}
    int size() const { return (int)neighbors_.size(); }
    BlockID target(int i) const { return neighbors_[i]; }
    void add_neighbor(const BlockID &b) { neighbors_.push_back(b); }
  };

  struct RegularContinuousLink : public Link
  {
    int dim_;
    ContinuousBounds core_, bounds_;
    std::vector<ContinuousBounds> nbr_bounds_;
    std::vector<Direction> dirs_, wraps_;
    RegularContinuousLink() : dim_(3) {}
    RegularContinuousLink(int dim, const ContinuousBounds &core, const ContinuousBounds &bounds)
        : dim_(dim), core_(core), bounds_(bounds) {}
    int dimension() const { return dim_; }
    const ContinuousBounds &core() const { return core_; }
    const ContinuousBounds &bounds() const { return bounds_; }
    const ContinuousBounds &bounds(int i) const { return nbr_bounds_[i]; }
    Direction direction(int i) const { return dirs_[i]; }
    Direction wrap(int i) const { return wraps_[i]; }
    void add_bounds(const ContinuousBounds &b) { nbr_bounds_.push_back(b); }
    void add_direction(const Direction &d) { dirs_.push_back(d); }
    void add_wrap(const Direction &d) { wraps_.push_back(d); }
  };

  inline void wrap_bounds(ContinuousBounds &b, const Direction &wrap_dir, const ContinuousBounds &domain)
  {
    for (int i = 0; i < 3; i++) {
      b.min[i] += wrap_dir[i] * (domain.max[i] - domain.min[i]);
      b.max[i] += wrap_dir[i] * (domain.max[i] - domain.min[i]);
    }
  }

  // distance from a point to a box (0 iff inside the closed box)
  template <class P>
  float distance(const ContinuousBounds &b, const P &p)
  {
    float res = 0;
    for (int i = 0; i < 3; i++) {
      float diff = 0, d;
      d = b.min[i] - p[i]; if (d > diff) diff = d;
      d = p[i] - b.max[i]; if (d > diff) diff = d;
      res += diff * diff;
    }
    return res;
  }

  template <class P, class OutIter>
  void in(const RegularContinuousLink &link, const P &p, OutIter out, const ContinuousBounds &domain)
  {
    for (int n = 0; n < link.size(); n++) {
      ContinuousBounds nb = link.bounds(n);
      wrap_bounds(nb, link.wrap(n), domain);
      if (distance(nb, p) == 0) *out++ = n;
    }
  }

  // ---- mpi wrapper ----
  namespace mpi
  {
    struct communicator
    {
      MPI_Comm c;
      communicator(MPI_Comm c_ = MPI_COMM_WORLD) : c(c_) {}
      operator MPI_Comm() const { return c; }
      int rank() const { return 0; }
      int size() const { return 1; }
    };
    struct environment { environment(int, char **) {} environment() {} };
  }

  // ---- assigners ----
  struct Assigner
  {
    int size_, nblocks_;
    Assigner(int size, int nblocks) : size_(size), nblocks_(nblocks) {}
    virtual ~Assigner() {}
    int nblocks() const { return nblocks_; }
    virtual int rank(int) const { return 0; }
  };
  struct StaticAssigner : public Assigner { using Assigner::Assigner; };
  struct ContiguousAssigner : public StaticAssigner { using StaticAssigner::StaticAssigner; };
  struct RoundRobinAssigner : public StaticAssigner { using StaticAssigner::StaticAssigner; };

  // ---- master ----
  class Master
  {
  public:
    typedef std::map<int, MemoryBuffer> Queues;   // keyed by gid, ascending

    struct ProxyWithLink
    {
      Master *master_;
      int lid_, gid_;
      Link *link_;
      ProxyWithLink(Master *m, int lid, int gid, Link *l) : master_(m), lid_(lid), gid_(gid), link_(l) {}
      int gid() const { return gid_; }
      Link *link() const { return link_; }

      template <class T>
      void enqueue(const BlockID &to, const T &x) const { diy::save(master_->outgoing_[lid_][to.gid], x); }

      void incoming(std::vector<int> &v) const
      {
        for (Queues::iterator it = master_->incoming_[lid_].begin(); it != master_->incoming_[lid_].end(); ++it)
          v.push_back(it->first);
      }
      MemoryBuffer &incoming(int from) const { return master_->incoming_[lid_][from]; }

      template <class T>
      void dequeue(int from, T *x, size_t n) const
      {
        MemoryBuffer &bb = master_->incoming_[lid_][from];
        diy::load(bb, x, n);
      }
      template <class T>
      void dequeue(int from, T &x) const { diy::load(master_->incoming_[lid_][from], x); }
    };

    Master() {}
    Master(mpi::communicator c, int = 1, int = -1) : comm_(c) {}
    ~Master() { for (size_t i = 0; i < links_.size(); i++) delete links_[i]; }

    int add(int gid, void *b, Link *l)
    {
      blocks_.push_back(b); gids_.push_back(gid); links_.push_back(l);
      lids_[gid] = (int)blocks_.size() - 1;
      outgoing_.push_back(Queues()); incoming_.push_back(Queues());
      return (int)blocks_.size() - 1;
    }
    unsigned size() const { return (unsigned)blocks_.size(); }
    int lid(int gid) const { return lids_.find(gid)->second; }
    int gid(int lid) const { return gids_[lid]; }
    template <class B> B *block(int i) const { return static_cast<B *>(blocks_[i]); }
    const mpi::communicator &communicator() const { return comm_; }

    template <class F>
    void foreach(const F &f)
    {
      for (size_t i = 0; i < blocks_.size(); i++) {
        ProxyWithLink cp(this, (int)i, gids_[i], links_[i]);
        call(f, blocks_[i], cp);
      }
    }

    void exchange()
    {
      for (size_t i = 0; i < incoming_.size(); i++) incoming_[i].clear();
      for (size_t i = 0; i < outgoing_.size(); i++) {
        for (Queues::iterator it = outgoing_[i].begin(); it != outgoing_[i].end(); ++it) {
          std::map<int, int>::iterator l = lids_.find(it->first);
          if (l == lids_.end()) continue;
          MemoryBuffer &dst = incoming_[l->second][gids_[i]];
          dst.buffer.swap(it->second.buffer);
          dst.position = 0;
        }
        outgoing_[i].clear();
      }
    }

  private:
    // deduce the block type from the callable's first parameter
    template <class F, class B>
    static void call_impl(const F &f, void *b, const ProxyWithLink &cp, void (F::*)(B *, const ProxyWithLink &) const)
    {
      f(static_cast<B *>(b), cp);
    }
    template <class F>
    static void call(const F &f, void *b, const ProxyWithLink &cp) { call_impl(f, b, cp, &F::operator()); }

    mpi::communicator comm_;
    std::vector<void *> blocks_;
    std::vector<int> gids_;
    std::vector<Link *> links_;
    std::map<int, int> lids_;
    std::vector<Queues> outgoing_, incoming_;
    friend struct ProxyWithLink;
  };
}

#endif
