// C-ABI entry points of liblcx_b200.so (declared in include/lcx_b200.h): engine life cycle, buffer management,
// host<->device transfers, and thin exception-to-status wrappers around the passes implemented in the other
// translation units.
#include "lcx_engine.cuh"

#include <cstdlib>

#include <map>
#include <memory>
#include <sstream>

namespace
{
  thread_local std::string g_last_error;

  template <class F>
  int guarded(F f)
  {
    try { f(); return 0; }
    catch (const std::exception &ex) { g_last_error = ex.what(); return 1; }
    catch (...) { g_last_error = "unknown error"; return 2; }
  }

  // consumer = the call may read or write cell fields on the engine's stream: order it after pending uploads
  // keep_pending = the call works on cell fields only (or is the condensation entry point, which consumes a pending
  // gather-on-read re-layout itself): everything else completes it first
  void use_device(lcx_engine *e, bool consumer = true, unsigned keep_pending = 0u)
  {
    LCX_CUDA(cudaSetDevice(e->device));
    if (e->pending & ~keep_pending) lcx::finish_pending(e, e->pending & ~keep_pending);
    if (!consumer) return;
    if (e->scalars_pending)
    {
      LCX_CUDA(cudaStreamWaitEvent(e->stream, e->scalars_ready, 0));
      e->scalars_pending = false;
    }
    // read-backs queued on the third stream: whatever comes next may overwrite what they read - except inside a chunked step,
    // where the next chunk's kernels touch other cells only (the batch ends with lcx_sync)
    if (e->d2h_open && e->win_end == 0)
    {
      LCX_CUDA(cudaEventRecord(e->d2h_mark, e->d2h_stream));
      LCX_CUDA(cudaStreamWaitEvent(e->stream, e->d2h_mark, 0));
      e->d2h_open = false;
    }
    e->tail_is_gather = false;
  }

  struct field_ref { lcx::real_t *p; size_t n; };

  field_ref field_of(lcx_engine *e, int field)
  {
    switch (field)
    {
      case LCX_F_TH:        return {e->th.p, e->th.n};
      case LCX_F_RV:        return {e->rv.p, e->rv.n};
      case LCX_F_RHOD:      return {e->rhod.p, e->rhod.n};
      case LCX_F_P:         return {e->p.p, e->p.n};
      case LCX_F_COURANT_X: return {e->courant_x.p, e->courant_x.n};
      case LCX_F_COURANT_Y: return {e->courant_y.p, e->courant_y.n};
      case LCX_F_COURANT_Z: return {e->courant_z.p, e->courant_z.n};
      case LCX_F_T:         return {e->T.p, e->T.n};
      case LCX_F_RH:        return {e->RH.p, e->RH.n};
      case LCX_F_ETA:       return {e->eta.p, e->eta.n};
      case LCX_F_DV:        return {e->dv.p, e->dv.n};
      case LCX_F_W_LS:      return {e->w_LS.p, e->w_LS.n};
      case LCX_F_MOM:       return {e->count_mom.p, e->count_mom.n};
      default: throw lcx::error("unknown field id " + std::to_string(field));
    }
  }

  // cell volumes, clipped by the Lagrangian domain: reference src/impl/initialization/particles_impl_init_grid.ipp:13-55
  __global__ void k_init_dv(lcx::grid_t g, lcx::real_t *dv)
  {
    using lcx::real_t;
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= g.n_cell) return;
    const int nz1 = max(1, g.nz), ny1 = max(1, g.ny);
    const int ic = int(c);
    const int i = (ic / nz1) / ny1, j = (ic / nz1) % ny1, k = ic % nz1;
    dv[c] = lcx::tmax(real_t(0),
      (lcx::tmin((i + 1) * g.dx, g.x1) - lcx::tmax(i * g.dx, g.x0)) *
      (lcx::tmin((j + 1) * g.dy, g.y1) - lcx::tmax(j * g.dy, g.y0)) *
      (lcx::tmin((k + 1) * g.dz, g.z1) - lcx::tmax(k * g.dz, g.z0)));
  }

  __global__ void k_fill_tail(size_t first, size_t count, lcx::real_t *vt, lcx::real_t *rc2, uint32_t *sid, uint32_t *ijk, const uint32_t *ijk_src)
  {
    const size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= count) return;
    vt[first + t] = lcx::real_t(-1);          // "invalid": resize value of vt (particles_impl.ipp:446, hskpng_resize.ipp:14-20)
    if (rc2) rc2[first + t] = lcx::real_t(-1);   // the same for rc2 (particles_impl.ipp:490)
    sid[first + t] = uint32_t(first + t);
    ijk[first + t] = ijk_src ? ijk_src[t] : 0u;
  }
}

lcx_engine::~lcx_engine()
{
  cudaSetDevice(device);
  if (stream) cudaStreamSynchronize(stream);
  sd[0].release(); sd[1].release();
  key[0].release(); key[1].release(); val[0].release(); val[1].release(); un.release(); flag.release();
  u01.release(); n_filtered.release(); tmp_real.release(); perm.release(); pending_perm.release();
  th.release(); rv.release(); rhod.release(); p.release(); T.release(); RH.release(); eta.release(); dv.release();
  lambda_D.release(); lambda_K.release(); sstp_tmp_rv.release(); sstp_tmp_th.release(); sstp_tmp_rh.release();
  drw_mom3.release(); rw_mom3.release(); count_mom.release(); mom_partial.release();
  courant_x.release(); courant_y.release(); courant_z.release(); w_LS.release(); cell_off.release(); cell_off_new.release(); arr_off.release(); mv_scan.release();
  vt0.release(); eff.release(); hist.release(); scan_tmp.release();
  for (int s = 0; s < 2; ++s)
  {
    for (int d = 0; d < 2; ++d) { mig_key[s][d].release(); mig_val[s][d].release(); }
    if (remote[s].ipc && remote[s].base) cudaIpcCloseMemHandle(remote[s].base);
    inbox[s].release();
  }
  if (ev_put) cudaEventDestroy(ev_put);
  scalars.release(); red_partial.release(); cell_tmp4.release(); big_cells.release();
  for (auto &r : prof) { cudaEventDestroy(r.t0); cudaEventDestroy(r.t1); }
  if (timer0) { cudaEventDestroy(timer0); cudaEventDestroy(timer1); }
  if (h_scalars) cudaFreeHost(h_scalars);
  if (scalars_ready) cudaEventDestroy(scalars_ready);
  if (pre_gather) cudaEventDestroy(pre_gather);
  if (courant_ready) cudaEventDestroy(courant_ready);
  if (main_mark) cudaEventDestroy(main_mark);
  if (copy_stream) cudaStreamDestroy(copy_stream);
  if (d2h_stream) cudaStreamDestroy(d2h_stream);
  if (win_stream) { if (stream == win_stream) stream = win_home; cudaStreamDestroy(win_stream); }
  if (win_join) cudaEventDestroy(win_join);
  if (d2h_mark) cudaEventDestroy(d2h_mark);
  if (stream) cudaStreamDestroy(stream);
}

namespace lcx
{
  void wait_courant(lcx_engine *e)
  {
    if (!e->courant_pending) return;
    LCX_CUDA(cudaStreamWaitEvent(e->stream, e->courant_ready, 0));
    e->courant_pending = false;
  }
}

extern "C" {

const char *lcx_last_error(void) { return g_last_error.c_str(); }
const char *lcx_version(void) { return "lcx_b200 0.1 (sm_100a)"; }

int lcx_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int lcx_create(const lcx_config *cfg, lcx_engine **out)
{
  *out = nullptr;
  return guarded([&] {
    using namespace lcx;
    if (cfg->real_bytes != int(sizeof(real_t))) throw error("lcx_create: this engine computes in " + std::string(sizeof(real_t) == 8 ? "double" : "single") + " precision; real_bytes does not match");
    if (lcx_device_count() == 0) throw error("lcx_create: no CUDA device is available - the B200 back-end has no CPU fallback");
    if (cfg->n_sd_max == 0) throw error("lcx_create: n_sd_max must be positive");
    if (cfg->n_sd_max >= 0xfffffff0ull) throw error("lcx_create: n_sd_max per slab must be below 2^32");
    std::unique_ptr<lcx_engine> e(new lcx_engine);
    e->cfg = *cfg;
    int dev = cfg->device;
    if (dev < 0) LCX_CUDA(cudaGetDevice(&dev));
    e->device = dev;
    LCX_CUDA(cudaSetDevice(dev));
    LCX_CUDA(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    LCX_CUDA(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
    LCX_CUDA(cudaStreamCreateWithFlags(&e->d2h_stream, cudaStreamNonBlocking));
    LCX_CUDA(cudaStreamCreateWithFlags(&e->win_stream, cudaStreamNonBlocking));
    LCX_CUDA(cudaEventCreateWithFlags(&e->win_join, cudaEventDisableTiming));
    LCX_CUDA(cudaEventCreateWithFlags(&e->d2h_mark, cudaEventDisableTiming));
    LCX_CUDA(cudaEventCreateWithFlags(&e->courant_ready, cudaEventDisableTiming));
    LCX_CUDA(cudaEventCreateWithFlags(&e->main_mark, cudaEventDisableTiming));
    LCX_CUDA(cudaEventCreateWithFlags(&e->scalars_ready, cudaEventDisableTiming));
    LCX_CUDA(cudaEventCreateWithFlags(&e->pre_gather, cudaEventDisableTiming));

    grid_t &g = e->grid;
    g.nx = cfg->nx; g.ny = cfg->ny; g.nz = cfg->nz;
    g.n_dims = (g.nx ? 1 : 0) + (g.ny ? 1 : 0) + (g.nz ? 1 : 0);
    if (g.ny && !(g.nx && g.nz)) throw error("lcx_create: a y dimension requires x and z (3-D is xyz, 2-D is xz, 1-D is x)");
    if (g.nz && !g.nx) throw error("lcx_create: a z dimension requires x");
    g.dx = cfg->dx; g.dy = cfg->dy; g.dz = cfg->dz;
    g.x0 = cfg->x0; g.y0 = cfg->y0; g.z0 = cfg->z0; g.x1 = cfg->x1; g.y1 = cfg->y1; g.z1 = cfg->z1;
    const size_t nx1 = g.nx ? g.nx : 1, ny1 = g.ny ? g.ny : 1, nz1 = g.nz ? g.nz : 1;
    const size_t n_cell = nx1 * ny1 * nz1;
    if (n_cell >= 0x7ffffff0ull) throw error("lcx_create: too many cells for one slab");
    g.n_cell = idx_t(n_cell);
    g.halo_size = (cfg->adve_scheme == AS_PRED_CORR) ? 2 : 0;
    {
      // spare bits of the last radix digit (lcx_post_copy sorts on 8-bit digits); LCX_SIZE_CLASS_BITS=0 switches the sub-key off
      int bits = 0; for (uint64_t v = n_cell; v; v >>= 1) ++bits;
      const int spare = (8 - bits % 8) % 8;
      const char *env = std::getenv("LCX_SIZE_CLASS_BITS");
      const int want = env ? std::atoi(env) : 0;      // off by default: the movers-only re-layout keeps arrival order inside a cell
      g.class_bits = spare < want ? spare : want;
      if (g.class_bits < 0) g.class_bits = 0;
      if (g.class_bits > 3) g.class_bits = 3;
    }
    g.halo_x = idx_t(g.n_dims == 1 ? g.halo_size : g.n_dims == 2 ? g.halo_size * g.nz : g.halo_size * g.nz * g.ny);

    const size_t cap = e->cap = size_t(cfg->n_sd_max);
    // gather-on-read re-layout (lcx_engine.cuh): on unless LCX_LAZY_GATHER=0 (measured on the cfg4 slab: 16.6 -> 14.6 ms per step)
    { const char *lz = std::getenv("LCX_LAZY_GATHER"); e->lazy_gather = !(lz && lz[0] == '0'); }
    e->sd[0].alloc(cap, g.nx, g.ny, g.nz);
    e->sd[1].alloc(cap, g.nx, g.ny, g.nz);
    if (cfg->exact_sstp_cond && cfg->allow_sstp_cond)
      for (int b = 0; b < 2; ++b)
      {
        e->sd[b].alloc_pp(cap, cfg->const_p != 0);
        if (cfg->sstp_cond_act > 1) e->sd[b].rc2.alloc(cap);
      }
    for (int b = 0; b < 2; ++b) { e->key[b].alloc(cap); e->val[b].alloc(cap); }
    e->un.alloc(cap); e->flag.alloc(cap); e->u01.alloc(cap); e->n_filtered.alloc(cap); e->tmp_real.alloc(cap);

    e->th.alloc(n_cell); e->rv.alloc(n_cell); e->rhod.alloc(n_cell); e->p.alloc(n_cell); e->T.alloc(n_cell);
    e->RH.alloc(n_cell); e->eta.alloc(n_cell); e->dv.alloc(n_cell); e->lambda_D.alloc(n_cell); e->lambda_K.alloc(n_cell);
    e->drw_mom3.alloc(n_cell); e->rw_mom3.alloc(n_cell); e->count_mom.alloc(n_cell);
    if (cfg->terminal_velocity == lcx::VT_BEARD77 || cfg->terminal_velocity == lcx::VT_BEARD77FAST) e->cell_tmp4.alloc(size_t(n_cell) * 4);
    if (cfg->allow_sstp_cond) { e->sstp_tmp_rv.alloc(n_cell); e->sstp_tmp_th.alloc(n_cell); e->sstp_tmp_rh.alloc(n_cell); }
    e->big_cells.alloc(n_cell);
    e->cell_off.alloc(n_cell + 2); e->cell_off_new.alloc(n_cell + 2); e->arr_off.alloc(n_cell + 2); e->mv_scan.alloc(n_cell + 2);
    const size_t h = size_t(g.halo_size);
    switch (g.n_dims)   // staggered Courant fields with x-halo: init_sync.ipp:29-44
    {
      case 3:
        e->courant_x.alloc((g.nx + 2 * h + 1) * g.ny * g.nz);
        e->courant_y.alloc((g.nx + 2 * h) * (g.ny + 1) * g.nz);
        e->courant_z.alloc((g.nx + 2 * h) * g.ny * (g.nz + 1));
        break;
      case 2:
        e->courant_x.alloc((g.nx + 2 * h + 1) * g.nz);
        e->courant_z.alloc((g.nx + 2 * h) * (g.nz + 1));
        break;
      case 1:
        e->courant_x.alloc(g.nx + 2 * h + 1);
        break;
      default: break;
    }
    for (dbuf<real_t> *c : {&e->courant_x, &e->courant_y, &e->courant_z})
      if (c->n) LCX_CUDA(cudaMemsetAsync(c->p, 0, c->bytes(), e->stream));

    const size_t tiles = cap / 2048 + 2;
    e->hist.alloc(tiles * 256);
    e->scan_tmp.alloc((tiles * 256) / 1024 * 2 + cap / 1024 * 2 + 8192);   // per-level tile sums of the recursive scan

    if (cfg->bcond_lft == LCX_BCOND_DISTMEM || cfg->bcond_rgt == LCX_BCOND_DISTMEM)
    {
      // one x-column's worth of capacity: with |C_x| <= 1 no more than that can cross a face in one step
      e->mig_cap = cap / nx1 + 4096;
      if (e->mig_cap > cap) e->mig_cap = cap;
      lcx_migr_real_attrs(e.get(), &e->mig_n_real);
      for (int s = 0; s < 2; ++s)
      {
        for (int d = 0; d < 2; ++d) { e->mig_key[s][d].alloc(e->mig_cap); e->mig_val[s][d].alloc(e->mig_cap); }
        e->inbox[s].alloc(inbox_bytes(e->mig_cap, e->mig_n_real, lcx::halo_values(g)));
        LCX_CUDA(cudaMemsetAsync(e->inbox[s].p, 0, MIG_HDR_BYTES, e->stream));
      }
      LCX_CUDA(cudaEventCreateWithFlags(&e->ev_put, cudaEventDisableTiming));
    }

    e->scalars.alloc(1);
    LCX_CUDA(cudaMemsetAsync(e->scalars.p, 0, sizeof(dev_scalars), e->stream));
    LCX_CUDA(cudaMallocHost(reinterpret_cast<void **>(&e->h_scalars), sizeof(dev_scalars)));
    std::memset(e->h_scalars, 0, sizeof(dev_scalars));
    e->red_partial.alloc(cap / 256 * 4 + 4096);

    if (g.n_dims > 0) { k_init_dv<<<div_up(n_cell, 256), 256, 0, e->stream>>>(g, e->dv.p); LCX_CUDA(cudaGetLastError()); ++e->launches; }
    LCX_CUDA(cudaMemsetAsync(e->cell_off.p, 0, e->cell_off.bytes(), e->stream));
    LCX_CUDA(cudaStreamSynchronize(e->stream));
    *out = e.release();
  });
}

int lcx_destroy(lcx_engine *e) { return guarded([&] { delete e; }); }

int lcx_sync(lcx_engine *e)
{
  return guarded([&] {
    use_device(e, true, 3u);
    LCX_CUDA(cudaStreamSynchronize(e->stream));
    if (e->upload_batch_open) LCX_CUDA(cudaStreamSynchronize(e->copy_stream));   // the data stay "pending" for the engine's stream
    e->upload_batch_open = false;
    if (e->d2h_open) LCX_CUDA(cudaStreamSynchronize(e->d2h_stream));
    e->d2h_open = false;
  });
}

int lcx_set_cell_window(lcx_engine *e, int64_t c_begin, int64_t c_end)
{
  return guarded([&] {
    LCX_CUDA(cudaSetDevice(e->device));
    if (c_begin == 0 && c_end == 0)
    {
      if (e->win_end == 0) return;
      // back to the engine's own stream, behind the chunks that ran on the second one
      if (e->stream == e->win_stream)
      {
        LCX_CUDA(cudaEventRecord(e->win_join, e->win_stream));
        e->stream = e->win_home;
        LCX_CUDA(cudaStreamWaitEvent(e->stream, e->win_join, 0));
      }
      else if (e->win_count > 1)
      {
        LCX_CUDA(cudaEventRecord(e->win_join, e->win_stream));
        LCX_CUDA(cudaStreamWaitEvent(e->stream, e->win_join, 0));
      }
      e->win_begin = e->win_end = 0;
      e->win_count = 0;
      return;
    }
    if (c_begin < 0 || c_end <= c_begin || c_end > int64_t(e->grid.n_cell)) throw lcx::error("lcx_set_cell_window: bad range");
    // Consecutive chunks alternate between two streams: they work on disjoint cells and SDs, so chunk k + 1 (as soon as its
    // fields have arrived) fills the SMs that chunk k's last wave of CTAs leaves idle instead of waiting for the kernel to end
    if (e->win_end == 0)
    {
      e->win_home = e->stream;
      LCX_CUDA(cudaEventRecord(e->win_join, e->stream));           // fork: the second stream starts behind what is queued so far
      LCX_CUDA(cudaStreamWaitEvent(e->win_stream, e->win_join, 0));
      e->win_count = 0;
    }
    e->stream = (e->win_count & 1) ? e->win_stream : e->win_home;
    ++e->win_count;
    e->win_begin = lcx::idx_t(c_begin); e->win_end = lcx::idx_t(c_end);
  });
}

int lcx_cond_granule(lcx_engine *e, int64_t *cells) { return guarded([&] { *cells = lcx::cond_granule(e); }); }

void *lcx_stream(lcx_engine *e) { return e->stream; }

int lcx_field_size(lcx_engine *e, int field, int64_t *count) { return guarded([&] { *count = int64_t(field_of(e, field).n); }); }

int lcx_cells_set(lcx_engine *e, int field, const void *src, int64_t count, int src_on_device)
{
  return guarded([&] {
    use_device(e, true, 3u);
    if (field == LCX_F_W_LS && e->w_LS.n != size_t(count)) e->w_LS.alloc(size_t(count));
    const field_ref f = field_of(e, field);
    lcx::wait_courant(e);
    if (size_t(count) != f.n) throw lcx::error("lcx_cells_set: field " + std::to_string(field) + " holds " + std::to_string(f.n) + " values, got " + std::to_string(count));
    LCX_CUDA(cudaMemcpyAsync(f.p, src, f.n * sizeof(lcx::real_t), src_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, e->stream));
    if (!src_on_device) LCX_CUDA(cudaStreamSynchronize(e->stream));   // the caller may reuse its staging buffer
  });
}

int lcx_cells_get(lcx_engine *e, int field, void *dst, int64_t count)
{
  return guarded([&] {
    use_device(e, true, 3u);
    const field_ref f = field_of(e, field);
    lcx::wait_courant(e);
    if (size_t(count) > f.n) throw lcx::error("lcx_cells_get: requested more values than the field holds");
    LCX_CUDA(cudaMemcpyAsync(dst, f.p, size_t(count) * sizeof(lcx::real_t), cudaMemcpyDeviceToHost, e->stream));
    LCX_CUDA(cudaStreamSynchronize(e->stream));
  });
}

int lcx_cells_set_part(lcx_engine *e, int field, int64_t offset, const void *src, int64_t count)
{
  return guarded([&] {
    const bool after_relayout = e->tail_is_gather;      // nothing but the gather has been queued since the last re-layout
    use_device(e, /*consumer=*/false, /*keep_pending=*/3u);
    const field_ref f = field_of(e, field);
    if (offset < 0 || count < 0 || size_t(offset + count) > f.n) throw lcx::error("lcx_cells_set_part: range outside field " + std::to_string(field));
    const bool courant = field == LCX_F_COURANT_X || field == LCX_F_COURANT_Y || field == LCX_F_COURANT_Z;
    if (!e->upload_batch_open)      // first piece of a batch: what is queued on the engine's stream may still read the old fields
    {
      if (after_relayout) LCX_CUDA(cudaStreamWaitEvent(e->copy_stream, e->pre_gather, 0));     // ... except the gather
      else
      {
        LCX_CUDA(cudaEventRecord(e->main_mark, e->stream));
        LCX_CUDA(cudaStreamWaitEvent(e->copy_stream, e->main_mark, 0));
      }
      e->upload_batch_open = true;
    }
    LCX_CUDA(cudaMemcpyAsync(static_cast<lcx::real_t *>(f.p) + offset, src, size_t(count) * sizeof(lcx::real_t), cudaMemcpyDefault, e->copy_stream));
    if (courant) { LCX_CUDA(cudaEventRecord(e->courant_ready, e->copy_stream)); e->courant_pending = true; }
    else         { LCX_CUDA(cudaEventRecord(e->scalars_ready, e->copy_stream)); e->scalars_pending = true; }
  });
}

int lcx_cells_get_part(lcx_engine *e, int field, int64_t offset, void *dst, int64_t count)
{
  return guarded([&] {
    use_device(e, true, 3u);
    const field_ref f = field_of(e, field);
    if (offset < 0 || count < 0 || size_t(offset + count) > f.n) throw lcx::error("lcx_cells_get_part: range outside field " + std::to_string(field));
    // on a third stream, behind what is queued on the engine's stream so far: in a chunked step the next chunk's kernels need not
    // wait for this chunk's read-back (lcx_sync ends the batch)
    LCX_CUDA(cudaEventRecord(e->main_mark, e->stream));
    LCX_CUDA(cudaStreamWaitEvent(e->d2h_stream, e->main_mark, 0));
    LCX_CUDA(cudaMemcpyAsync(dst, static_cast<const lcx::real_t *>(f.p) + offset, size_t(count) * sizeof(lcx::real_t), cudaMemcpyDefault, e->d2h_stream));
    e->d2h_open = true;
  });
}

int lcx_pointer_on_device(const void *p, int *on_device)
{
  *on_device = 0;
  return guarded([&] {
    if (!p) return;
    cudaPointerAttributes at;
    const cudaError_t st = cudaPointerGetAttributes(&at, p);
    if (st != cudaSuccess) { cudaGetLastError(); return; }      // plain pageable host memory on older drivers
    *on_device = (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) ? 1 : 0;
  });
}

int lcx_copy_to_host(void *dst_host, const void *src_any, size_t bytes)
{ return guarded([&] { if (bytes) LCX_CUDA(cudaMemcpy(dst_host, src_any, bytes, cudaMemcpyDefault)); }); }

int lcx_host_alloc(size_t bytes, void **out)
{
  *out = nullptr;
  return guarded([&] { LCX_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault)); });
}

int lcx_host_free(void *p) { return guarded([&] { if (p) LCX_CUDA(cudaFreeHost(p)); }); }

int lcx_set_vt0_table(lcx_engine *e, const void *table, int n)
{
  return guarded([&] {
    use_device(e);
    if (n != lcx::VT0_N_BIN) throw lcx::error("lcx_set_vt0_table: expected " + std::to_string(int(lcx::VT0_N_BIN)) + " bins");
    e->vt0.alloc(size_t(n));
    LCX_CUDA(cudaMemcpy(e->vt0.p, table, size_t(n) * sizeof(lcx::real_t), cudaMemcpyHostToDevice));
  });
}

int lcx_set_efficiencies(lcx_engine *e, const void *table, int64_t n)
{
  return guarded([&] {
    use_device(e);
    e->eff.alloc(size_t(n));
    LCX_CUDA(cudaMemcpy(e->eff.p, table, size_t(n) * sizeof(lcx::real_t), cudaMemcpyHostToDevice));
  });
}

int lcx_sd_append(lcx_engine *e, int64_t count, const uint64_t *n, const void *rd3, const void *rw2, const void *kpa,
                  const void *x, const void *y, const void *z, const uint32_t *ijk)
{
  return guarded([&] {
    using namespace lcx;
    use_device(e);
    if (count <= 0) return;
    lcx::densify_sid(e);
    const size_t first = e->n_part, cnt = size_t(count);
    if (first + cnt > e->cap) throw error("n_sd_max (" + std::to_string(e->cap) + ") < n_part (" + std::to_string(first + cnt) + ")");
    sd_arrays &s = e->S();
    auto up = [&](void *dst, const void *src, size_t bytes) { if (dst && src) LCX_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, e->stream)); };
    up(s.n.p + first, n, cnt * sizeof(n_t));
    up(s.rd3.p + first, rd3, cnt * sizeof(real_t));
    up(s.rw2.p + first, rw2, cnt * sizeof(real_t));
    up(s.kpa.p + first, kpa, cnt * sizeof(real_t));
    if (e->grid.nx) up(s.x.p + first, x, cnt * sizeof(real_t));
    if (e->grid.ny) up(s.y.p + first, y, cnt * sizeof(real_t));
    if (e->grid.nz) up(s.z.p + first, z, cnt * sizeof(real_t));
    uint32_t *ijk_dev = nullptr;
    if (ijk) { ijk_dev = e->key[0].p; up(ijk_dev, ijk, cnt * sizeof(uint32_t)); }
    k_fill_tail<<<div_up(cnt, 256), 256, 0, e->stream>>>(first, cnt, s.vt.p, s.rc2.p, s.sid.p, s.ijk.p, ijk_dev);
    LCX_CUDA(cudaGetLastError()); ++e->launches;
    LCX_CUDA(cudaStreamSynchronize(e->stream));
    e->n_part = first + cnt;
    e->sid_hi = e->n_part;
    e->grouped = false;
    e->keys_ready = 0;
  });
}

int lcx_sd_append_sd_conc(lcx_engine *e, int64_t per_cell, double log_rd_min, double log_rd_max, double kappa, double RH_max,
                          uint64_t seed, uint32_t stream, uint64_t call, void *rd3_host)
{
  return guarded([&] {
    use_device(e);
    if (per_cell <= 0) return;
    lcx::sd_append_sd_conc(e, size_t(per_cell), lcx::real_t(log_rd_min), lcx::real_t(log_rd_max), lcx::real_t(kappa), lcx::real_t(RH_max),
                           seed, stream, call, static_cast<lcx::real_t *>(rd3_host));
  });
}

int lcx_sd_set_n(lcx_engine *e, int64_t first, int64_t count, const uint64_t *n)
{
  return guarded([&] {
    use_device(e);
    if (first < 0 || count < 0 || size_t(first + count) > e->n_part) throw lcx::error("lcx_sd_set_n: range outside the super-droplets");
    if (e->grouped) throw lcx::error("lcx_sd_set_n: call it right after lcx_sd_append_sd_conc (the re-layout has already moved the super-droplets)");
    LCX_CUDA(cudaMemcpyAsync(e->S().n.p + first, n, size_t(count) * sizeof(lcx::n_t), cudaMemcpyHostToDevice, e->stream));
    LCX_CUDA(cudaStreamSynchronize(e->stream));
  });
}

int lcx_set_dense_storage_index(lcx_engine *e, int always)
{ return guarded([&] { use_device(e); e->dense_always = always != 0; if (e->dense_always) lcx::densify_sid(e); }); }

int lcx_n_part(lcx_engine *e, int64_t *n_part) { *n_part = int64_t(e->n_part); return 0; }

int lcx_get_attr(lcx_engine *e, int attr, void *dst, int64_t cap, int64_t *n_out)
{
  return guarded([&] {
    use_device(e);
    *n_out = int64_t(e->n_part);
    if (e->n_part == 0) return;
    lcx::scatter_attr_by_sid(e, attr, e->tmp_real.p);
    const size_t cnt = size_t(cap) < e->n_part ? size_t(cap) : e->n_part;
    LCX_CUDA(cudaMemcpyAsync(dst, e->tmp_real.p, cnt * sizeof(lcx::real_t), cudaMemcpyDeviceToHost, e->stream));
    LCX_CUDA(cudaStreamSynchronize(e->stream));
  });
}

int lcx_get_attr_u64(lcx_engine *e, int attr, uint64_t *dst, int64_t cap, int64_t *n_out)
{
  return guarded([&] {
    use_device(e);
    *n_out = int64_t(e->n_part);
    if (e->n_part == 0) return;
    static_assert(sizeof(lcx::real_t) >= 4, "");
    const size_t cnt = size_t(cap) < e->n_part ? size_t(cap) : e->n_part;
    if (attr == LCX_A_N && sizeof(lcx::real_t) == sizeof(uint64_t))
    {
      // multiplicities travel as 64-bit integers: no round trip through real_t (exact beyond 2^53)
      lcx::scatter_n_by_sid(e, reinterpret_cast<uint64_t *>(e->tmp_real.p));
      LCX_CUDA(cudaMemcpyAsync(dst, e->tmp_real.p, cnt * sizeof(uint64_t), cudaMemcpyDeviceToHost, e->stream));
      LCX_CUDA(cudaStreamSynchronize(e->stream));
      return;
    }
    if (attr == LCX_A_N)
    {
      // single-precision engine: the real-sized scratch array is too small for 64-bit integers; a diagnostics path, so a
      // temporary allocation will do
      uint64_t *scratch = nullptr;
      LCX_CUDA(cudaMalloc(&scratch, e->n_part * sizeof(uint64_t)));
      lcx::scatter_n_by_sid(e, scratch);
      LCX_CUDA(cudaMemcpyAsync(dst, scratch, cnt * sizeof(uint64_t), cudaMemcpyDeviceToHost, e->stream));
      LCX_CUDA(cudaStreamSynchronize(e->stream));
      LCX_CUDA(cudaFree(scratch));
      return;
    }
    std::vector<lcx::real_t> tmp(e->n_part);
    lcx::scatter_attr_by_sid(e, attr, e->tmp_real.p);
    LCX_CUDA(cudaMemcpyAsync(tmp.data(), e->tmp_real.p, e->n_part * sizeof(lcx::real_t), cudaMemcpyDeviceToHost, e->stream));
    LCX_CUDA(cudaStreamSynchronize(e->stream));
    for (size_t i = 0; i < cnt; ++i) dst[i] = uint64_t(tmp[i]);
  });
}

int lcx_get_layout(lcx_engine *e, uint32_t *sid, uint32_t *ijk, int64_t cap, int64_t *n_out)
{
  return guarded([&] {
    use_device(e);
    *n_out = int64_t(e->n_part);
    const size_t cnt = size_t(cap) < e->n_part ? size_t(cap) : e->n_part;
    if (cnt == 0) return;
    if (sid) LCX_CUDA(cudaMemcpyAsync(sid, e->S().sid.p, cnt * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
    if (ijk) LCX_CUDA(cudaMemcpyAsync(ijk, e->S().ijk.p, cnt * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
    LCX_CUDA(cudaStreamSynchronize(e->stream));
  });
}

int lcx_hskpng_Tpr(lcx_engine *e) { return guarded([&] { use_device(e, true, 3u); lcx::hskpng_Tpr(e); }); }
int lcx_hskpng_mfp(lcx_engine *e) { return guarded([&] { use_device(e, true, 3u); lcx::hskpng_mfp(e); }); }
int lcx_hskpng_vterm(lcx_engine *e, int only_invalid) { return guarded([&] { use_device(e, true, lcx_engine::PENDING_XYZ); lcx::hskpng_vterm(e, only_invalid != 0); }); }
int lcx_sstp_percell_step(lcx_engine *e, int step, int sstp_cond, int var_rho)
{ return guarded([&] { use_device(e, true, 3u); lcx::sstp_percell_step(e, step, sstp_cond, var_rho != 0); }); }
int lcx_sstp_save(lcx_engine *e) { return guarded([&] { use_device(e, true, 3u); lcx::sstp_save(e); }); }

int lcx_set_cond_solver(int mode) { lcx::set_cond_solver(mode); return 0; }
int lcx_get_cond_solver(void) { return lcx::cond_solver(); }
int lcx_set_cond_layout(int cells_per_warp) { lcx::set_cond_layout(cells_per_warp); return 0; }
int lcx_get_cond_layout(void) { return lcx::cond_layout(); }
int lcx_set_cond_classed(int mode) { lcx::set_cond_classed(mode); return 0; }
int lcx_get_cond_classed(void) { return lcx::cond_classed(); }
int lcx_set_cond_staged(int on) { lcx::set_cond_staged(on); return 0; }
int lcx_get_cond_staged(void) { return lcx::cond_staged(); }

int lcx_cond_perparticle(lcx_engine *e, double dt, double RH_max, int sstp_cond, int mix)
{ return guarded([&] { use_device(e); lcx::cond_perparticle(e, dt, RH_max, sstp_cond, mix != 0); }); }

int lcx_cond_perparticle_adaptive(lcx_engine *e, double dt, double RH_max, int sstp_cond_max, int sstp_cond_act, double drw2_eps, double drw2_max)
{ return guarded([&] { use_device(e); lcx::cond_perparticle_adaptive(e, dt, RH_max, sstp_cond_max, sstp_cond_act, drw2_eps, drw2_max); }); }

int lcx_hskpng_rc2(lcx_engine *e) { return guarded([&] { use_device(e, true, lcx_engine::PENDING_XYZ); lcx::hskpng_rc2(e); }); }

int lcx_cond(lcx_engine *e, double dt_sub, double RH_max, int step, int sstp_cond)
{ return guarded([&] { use_device(e, true, 3u); lcx::cond(e, dt_sub, RH_max, step, sstp_cond); }); }

int lcx_coal(lcx_engine *e, double dt_sub, const lcx_rng *rng) { return guarded([&] { use_device(e, true, lcx_engine::PENDING_XYZ); lcx::coal(e, dt_sub, rng); }); }

int lcx_coal_flag(lcx_engine *e, int *increase_sstp_coal)
{
  return guarded([&] {
    use_device(e, true, 3u);      // device scalars only
    LCX_CUDA(cudaMemcpyAsync(e->h_scalars, e->scalars.p, sizeof(lcx::dev_scalars), cudaMemcpyDeviceToHost, e->stream));
    LCX_CUDA(cudaStreamSynchronize(e->stream));
    *increase_sstp_coal = int(e->h_scalars->increase_sstp_coal);
    if (*increase_sstp_coal) LCX_CUDA(cudaMemsetAsync(&e->scalars.p->increase_sstp_coal, 0, sizeof(unsigned int), e->stream));
  });
}

int lcx_coal_stats(lcx_engine *e, uint64_t *n_collisions, uint64_t *n_pairs_collided)
{
  return guarded([&] {
    use_device(e, true, 3u);      // device scalars only
    LCX_CUDA(cudaMemcpyAsync(e->h_scalars, e->scalars.p, sizeof(lcx::dev_scalars), cudaMemcpyDeviceToHost, e->stream));
    LCX_CUDA(cudaStreamSynchronize(e->stream));
    *n_collisions = e->h_scalars->n_collisions;
    *n_pairs_collided = e->h_scalars->n_pairs_collided;
  });
}

int lcx_transport(lcx_engine *e, const lcx_transport_opts *o) { return guarded([&] { use_device(e, true, lcx_engine::PENDING_XYZ); lcx::transport(e, o); }); }

int lcx_puddle(lcx_engine *e, double out[14])
{
  return guarded([&] {
    use_device(e, true, 3u);      // device scalars only
    LCX_CUDA(cudaMemcpyAsync(e->h_scalars, e->scalars.p, sizeof(lcx::dev_scalars), cudaMemcpyDeviceToHost, e->stream));
    LCX_CUDA(cudaStreamSynchronize(e->stream));
    for (int i = 0; i < 14; ++i) out[i] = 0;
    out[8] = e->h_scalars->puddle[0];    // outliq_vol
    out[9] = e->h_scalars->puddle[1];    // outdry_vol
    out[10] = e->h_scalars->puddle[3];   // outprtcl_num
    out[12] = e->h_scalars->puddle[2];   // outliq_num
  });
}

namespace
{
  struct ipc_blob { cudaIpcMemHandle_t handle; uint64_t cap; int32_t n_real; int32_t real_bytes; char pad[LCX_IPC_BLOB_BYTES - sizeof(cudaIpcMemHandle_t) - 16]; };
  static_assert(sizeof(ipc_blob) == LCX_IPC_BLOB_BYTES, "IPC blob layout");
}

int lcx_top_loss(lcx_engine *e, double out[2])
{
  return guarded([&] {
    use_device(e, true, 3u);      // device scalars only
    LCX_CUDA(cudaMemcpyAsync(e->h_scalars, e->scalars.p, sizeof(lcx::dev_scalars), cudaMemcpyDeviceToHost, e->stream));
    LCX_CUDA(cudaStreamSynchronize(e->stream));
    out[0] = e->h_scalars->puddle[4]; out[1] = e->h_scalars->puddle[5];
  });
}

int lcx_migr_connect(lcx_engine *e, int side, lcx_engine *nb)
{
  return guarded([&] {
    if (side < 0 || side > 1) throw lcx::error("lcx_migr_connect: bad side");
    if (!nb || !nb->inbox[side].p) throw lcx::error("lcx_migr_connect: the neighbour has no inbox on that side (not a distributed-memory slab)");
    if (nb->mig_n_real != e->mig_n_real) throw lcx::error("lcx_migr_connect: neighbours carry different attribute sets");
    LCX_CUDA(cudaSetDevice(e->device));
    if (nb->device != e->device)
    {
      int can = 0;
      LCX_CUDA(cudaDeviceCanAccessPeer(&can, e->device, nb->device));
      if (!can) throw lcx::error("lcx_migr_connect: device " + std::to_string(e->device) + " cannot access device " + std::to_string(nb->device) + " (no peer-to-peer path)");
      const cudaError_t st = cudaDeviceEnablePeerAccess(nb->device, 0);
      if (st == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError(); else LCX_CUDA(st);
    }
    e->remote[side].base = nb->inbox[side].p; e->remote[side].cap = nb->mig_cap; e->remote[side].ipc = false;
  });
}

int lcx_migr_ipc_export(lcx_engine *e, int side, void *blob)
{
  return guarded([&] {
    if (side < 0 || side > 1 || !e->inbox[side].p) throw lcx::error("lcx_migr_ipc_export: no inbox on that side");
    LCX_CUDA(cudaSetDevice(e->device));
    ipc_blob b;
    std::memset(&b, 0, sizeof(b));
    LCX_CUDA(cudaIpcGetMemHandle(&b.handle, e->inbox[side].p));
    b.cap = e->mig_cap; b.n_real = e->mig_n_real; b.real_bytes = int32_t(sizeof(lcx::real_t));
    std::memcpy(blob, &b, sizeof(b));
  });
}

int lcx_migr_ipc_connect(lcx_engine *e, int side, const void *blob)
{
  return guarded([&] {
    if (side < 0 || side > 1) throw lcx::error("lcx_migr_ipc_connect: bad side");
    ipc_blob b;
    std::memcpy(&b, blob, sizeof(b));
    if (b.n_real != e->mig_n_real || b.real_bytes != int32_t(sizeof(lcx::real_t))) throw lcx::error("lcx_migr_ipc_connect: neighbours carry different attribute sets");
    LCX_CUDA(cudaSetDevice(e->device));
    void *ptr = nullptr;
    LCX_CUDA(cudaIpcOpenMemHandle(&ptr, b.handle, cudaIpcMemLazyEnablePeerAccess));
    if (e->remote[side].ipc && e->remote[side].base) cudaIpcCloseMemHandle(e->remote[side].base);
    e->remote[side].base = static_cast<unsigned char *>(ptr); e->remote[side].cap = size_t(b.cap); e->remote[side].ipc = true;
  });
}

int lcx_migr_put(lcx_engine *e, int64_t *n_lft, int64_t *n_rgt) { return guarded([&] { use_device(e); lcx::migr_put(e, n_lft, n_rgt); }); }

int lcx_migr_take(lcx_engine *e, lcx_engine *rgt, lcx_engine *lft, int64_t *n_from_rgt, int64_t *n_from_lft)
{ return guarded([&] { use_device(e); lcx::migr_take(e, rgt, lft, n_from_rgt, n_from_lft); }); }

int lcx_halo_put(lcx_engine *e) { return guarded([&] { use_device(e, true, 3u); lcx::halo_put(e); }); }

int lcx_halo_take(lcx_engine *e) { return guarded([&] { use_device(e, true, 3u); lcx::halo_take(e); }); }

int lcx_migr_real_attrs(lcx_engine *e, int *count)
{
  *count = 4 + (e->grid.nx ? 1 : 0) + (e->grid.ny ? 1 : 0) + (e->grid.nz ? 1 : 0);
  if (e->cfg.exact_sstp_cond && e->cfg.allow_sstp_cond) *count += (e->cfg.const_p ? 4 : 3) + (e->cfg.sstp_cond_act > 1 ? 1 : 0);
  return 0;
}

int lcx_post_copy(lcx_engine *e, int rcyc, int keep_all) { return guarded([&] { use_device(e); lcx::post_copy(e, rcyc != 0, keep_all != 0); }); }

int lcx_moms_select(lcx_engine *e, int kind, int attr, double lo, double hi, int cons)
{ return guarded([&] { use_device(e); lcx::moms_select(e, kind, attr, lo, hi, cons != 0); }); }

int lcx_moms_calc(lcx_engine *e, int attr, double power, int specific)
{
  return guarded([&] {
    use_device(e);
    if (!e->selected) throw lcx::error("moment requested before a selector (diag_all, diag_*_rng, ...)");
    lcx::cell_moment(e, e->n_filtered.p, lcx::attr_ptr(e, attr), power, specific != 0, e->count_mom.p);
  });
}

int lcx_diag_sd_conc(lcx_engine *e) { return guarded([&] { use_device(e); lcx::diag_sd_conc(e); }); }

int lcx_diag_cell_field(lcx_engine *e, int field)
{
  return guarded([&] {
    use_device(e);
    const field_ref f = field_of(e, field);
    LCX_CUDA(cudaMemcpyAsync(e->count_mom.p, f.p, size_t(e->grid.n_cell) * sizeof(lcx::real_t), cudaMemcpyDeviceToDevice, e->stream));
  });
}

int lcx_diag_precip_rate(lcx_engine *e) { return guarded([&] { use_device(e); lcx::diag_precip_rate(e); }); }
int lcx_diag_mass_dens(lcx_engine *e, int attr, double rad, double sig0, double xp)
{ return guarded([&] { use_device(e); lcx::diag_mass_dens(e, attr, rad, sig0, xp); }); }
int lcx_diag_vel_div(lcx_engine *e, double dt) { return guarded([&] { use_device(e); lcx::diag_vel_div(e, dt); }); }
int lcx_diag_max_rw(lcx_engine *e) { return guarded([&] { use_device(e); lcx::diag_max_rw(e); }); }

int lcx_outbuf(lcx_engine *e, void *dst, int64_t count) { return lcx_cells_get(e, LCX_F_MOM, dst, count); }

int lcx_timer_start(lcx_engine *e)
{
  return guarded([&] {
    use_device(e, true, 3u);
    if (!e->timer0) { LCX_CUDA(cudaEventCreate(&e->timer0)); LCX_CUDA(cudaEventCreate(&e->timer1)); }
    LCX_CUDA(cudaEventRecord(e->timer0, e->stream));
  });
}

int lcx_timer_stop(lcx_engine *e, float *ms)
{
  return guarded([&] {
    use_device(e, true, 3u);
    if (!e->timer0) throw lcx::error("lcx_timer_stop without lcx_timer_start");
    LCX_CUDA(cudaEventRecord(e->timer1, e->stream));
    LCX_CUDA(cudaEventSynchronize(e->timer1));
    LCX_CUDA(cudaEventElapsedTime(ms, e->timer0, e->timer1));
  });
}

int lcx_profile_enable(lcx_engine *e, int on)
{
  return guarded([&] {
    use_device(e, true, 3u);
    LCX_CUDA(cudaStreamSynchronize(e->stream));
    for (auto &r : e->prof) { cudaEventDestroy(r.t0); cudaEventDestroy(r.t1); }
    e->prof.clear();
    e->profiling = on != 0;
  });
}

int lcx_profile_report(lcx_engine *e, char *buf, int64_t size)
{
  return guarded([&] {
    use_device(e, true, 3u);
    LCX_CUDA(cudaStreamSynchronize(e->stream));
    std::map<std::string, std::pair<uint64_t, double>> acc;
    for (auto &r : e->prof)
    {
      float ms = 0;
      LCX_CUDA(cudaEventElapsedTime(&ms, r.t0, r.t1));
      auto &a = acc[r.name];
      a.first += 1; a.second += ms;
    }
    std::ostringstream os;
    for (auto &kv : acc) os << kv.first << " " << kv.second.first << " " << kv.second.second << "\n";
    const std::string s = os.str();
    if (size > 0) { std::strncpy(buf, s.c_str(), size_t(size) - 1); buf[size - 1] = 0; }
  });
}

int lcx_launch_count(lcx_engine *e, uint64_t *launches) { *launches = e->launches; return 0; }

int lcx_cell_stats(lcx_engine *e, int64_t *n_cell, int64_t *max_count)
{ *n_cell = int64_t(e->grid.n_cell); *max_count = int64_t(e->max_count); return 0; }

}  // extern "C"
