// Super-droplet transport in one sweep: advection on the Arakawa-C Courant fields (implicit / explicit Euler /
// predictor-corrector), sedimentation, large-scale subsidence, boundary conditions with precipitation
// accounting, and detection of SDs that left the x-slab.  Plus the x-slab migration pack / unpack kernels.
//
// Reference: src/impl/advection/particles_impl_adve.ipp:27-304, src/impl/sedimentation/particles_impl_sedi.ipp:13-24,
//            src/impl/subsidence/particles_impl_subs.ipp:13-25, src/impl/boundary_conditions/particles_impl_bcnd.ipp:99-368,
//            src/impl/initialization/particles_impl_init_grid.ipp:93-155 (face-index tables, evaluated arithmetically here),
//            src/impl/distributed_memory/particles_impl_pack.ipp:15-121, particles_impl_unpack.ipp:15-145,
//            src/impl_multi_gpu/particles_multi_gpu_impl_step_async_and_copy.ipp:100-140.
//
// The reference spends ~20 Thrust passes (plus 4 device-wide reductions for the puddle) on this; here every SD is
// read once (x, y, z, vt, ijk, n: 48 B) and written once (x, y, z, n: 32 B).  Arithmetic order follows the reference
// so positions agree bit for bit (compile with -fmad=false).
#include "lcx_engine.cuh"

#include <cstdlib>

namespace lcx
{
  namespace
  {
    constexpr int TPB = 256;

    struct tr_params
    {
      grid_t g;
      int adve, sedi, subs, scheme;
      real_t dt;
      int open_side_walls, periodic_topbot, bcond_lft, bcond_rgt;
      // per-launch constants the host works out once instead of every thread: 1/dx.. (correctly rounded, as the kernel's own
      // real_t(1) / dx was), 1/(x1 - x0).. for the periodic wrap, and the magic numbers of the divisions by nz and ny
      real_t inv_dx, inv_dy, inv_dz, inv_Lx, inv_Ly, inv_Lz;
      uint32_t nz_m, nz_s1, nz_s2, ny_m, ny_s1, ny_s2;
    };

    // n / d for any 32-bit n by multiplication (Granlund & Montgomery 1994, fig. 4.1): m, s1, s2 from fastdiv_setup(d)
    __device__ __forceinline__ uint32_t fastdiv(uint32_t n, uint32_t m, uint32_t s1, uint32_t s2)
    {
      const uint32_t t = __umulhi(m, n);
      return (t + ((n - t) >> s1)) >> s2;
    }
    void fastdiv_setup(uint32_t d, uint32_t &m, uint32_t &s1, uint32_t &s2)
    {
      uint32_t l = 0;
      while ((uint64_t(1) << l) < d) ++l;                       // ceil(log2 d)
      m = uint32_t(((uint64_t(1) << 32) * ((uint64_t(1) << l) - d)) / d + 1);
      s1 = l < 1 ? l : 1; s2 = l < 1 ? 0 : l - 1;
    }

    // fmod(t, L) for t >= 0, L > 0 and a small quotient, without the library's iterative reduction.  fmod is exact in IEEE
    // arithmetic (the remainder is always representable), and t - q L evaluated by ONE fma is exact whenever it is
    // representable, i.e. as soon as q is the true integer quotient; a quotient off by one (rounding of t / L) is
    // detected by the sign / size of the remainder and the fma is redone.  Bit-identical to fmod().
    __device__ __forceinline__ real_t fmod_small_quotient(real_t t, real_t L, real_t inv_L)
    {
      real_t q = floor(t * inv_L);       // a quotient that may be off by one costs no more than the rounded t / L was: see above
      real_t r = fma(-q, L, t);
      if (r < 0)       { q -= 1; r = fma(-q, L, t); }
      else if (r >= L) { q += 1; r = fma(-q, L, t); }
      return r;
    }
    __device__ __forceinline__ real_t periodic_wrap(real_t x, real_t a, real_t b, real_t inv_L)   // bcnd.ipp:99-110
    {
      const real_t L = b - a, t = (x - a) + 10 * L;
      // the reference assumes |displacement| < 10 domain lengths; outside that (or for NaN) fall back to the library
      const real_t r = (t >= 0 && t < 64 * L) ? fmod_small_quotient(t, L, inv_L) : fmod(t, L);
      return a + r;
    }

    __device__ __forceinline__ real_t step_impl(real_t x, idx_t i, real_t C_l, real_t C_r, real_t dx)   // adve.ipp:27-60
    { return (x + dx * (C_l - i * (C_r - C_l))) / (1 - (C_r - C_l)); }

    __device__ __forceinline__ real_t step_expl(real_t x, idx_t i, real_t C_l, real_t C_r, real_t dx, bool apply)   // adve.ipp:62-93
    { return apply * x + (C_r - C_l) * (x - dx * i) + dx * C_l; }

    // staggered-field indices of the two faces of (halo-extended) cell cp = (i * ny + j) * nz + k in each direction:
    // init_grid.ipp:93-155.  i and j come from the caller (it has them anyway): no integer division here.
    struct faces { idx_t xl, xr, yl, yr, zl, zr; };
    __device__ __forceinline__ faces faces_of(const grid_t &g, idx_t cp, idx_t i, idx_t j)
    {
      faces f;
      f.xl = cp;
      f.xr = cp + (g.n_dims == 3 ? idx_t(g.nz) * g.ny : idx_t(g.nz));
      f.yl = f.yr = f.zl = f.zr = 0;
      if (g.n_dims == 3)
      {
        f.yl = cp + i * g.nz;                  // cp / (ny nz) = i
        f.yr = f.yl + g.nz;
        f.zl = cp + g.ny * i + j;              // (cp - i ny nz) / nz = j
        f.zr = f.zl + 1;
      }
      else if (g.n_dims == 2)
      {
        f.zl = cp + i;                         // cp / nz = i
        f.zr = f.zl + 1;
      }
      return f;
    }

    // size_t(double(x) / dx) of hskpng_ijk.ipp:171 without the division: x * (1/dx) differs from the correctly rounded quotient
    // by a few ulp (< 1e-12 for any grid index < 2^20), so its integer part is the same unless the quotient sits within 1e-9
    // of an integer - only then (an SD numerically on a cell face) is the IEEE division evaluated.  Bit-identical result.
    __device__ __forceinline__ idx_t index_of(real_t x, real_t dx, real_t inv_dx)
    {
      const double q = double(x) * double(inv_dx);
      const int qi = __double2int_rd(q);                        // floor and conversion in one instruction (q < 2^20 below)
      const double fr = q - double(qi);
      if (fr > 1e-9 && fr < 1 - 1e-9 && q < 1048576.) return idx_t(qi);
      return idx_t(size_t(double(x) / double(dx)));
    }

    __device__ __forceinline__ idx_t cell_of(const tr_params &P, real_t x, real_t y, real_t z, idx_t &i, idx_t &j, idx_t &k)   // hskpng_ijk.ipp:159-200
    {
      const grid_t &g = P.g;
      i = g.nx ? index_of(x, g.dx, P.inv_dx) : 0;
      j = g.ny ? index_of(y, g.dy, P.inv_dy) : 0;
      k = g.nz ? index_of(z, g.dz, P.inv_dz) : 0;
      switch (g.n_dims)
      {
        case 1: return i;
        case 2: return i * g.nz + k;
        case 3: return i * (idx_t(g.nz) * g.ny) + j * g.nz + k;
        default: return 0;
      }
    }

#ifndef LCX_TR_MINB
#define LCX_TR_MINB 6      // latency-bound kernel: 40 registers, 48 resident warps; measured best of {3,4,5,6}
#endif
    // leavers through a slab face are listed on the fly (bcnd.ipp:160-172 does two copy_if sweeps): storage index as the later
    // sort key, physical index as the value; the lists are unordered (atomic slots) until lcx_migr_put sorts them
    struct mig_lists { const idx_t *sid; uint32_t *key[2], *val[2]; unsigned int *count[2]; unsigned cap; double *top_loss; };

    // LAZY: positions still lie in the previous layout (lcx_engine::PENDING_XYZ): read through the permutation, written in place
    template <bool LAZY>
    __global__ void __launch_bounds__(TPB, LCX_TR_MINB) k_transport(size_t n_part, tr_params P,
                                                      real_t *__restrict__ xs, real_t *__restrict__ ys, real_t *__restrict__ zs,
                                                      const real_t *__restrict__ vt, const idx_t *__restrict__ ijk, n_t *__restrict__ ns,
                                                      const real_t *__restrict__ rw2, const real_t *__restrict__ rd3,
                                                      const real_t *__restrict__ Cx, const real_t *__restrict__ Cy, const real_t *__restrict__ Cz,
                                                      const real_t *__restrict__ w_LS, double *__restrict__ partial,
                                                      uint32_t *__restrict__ key,
                                                      const uint32_t *__restrict__ perm, const real_t *__restrict__ xi, const real_t *__restrict__ yi,
                                                      const real_t *__restrict__ zi, mig_lists M)
    {
      __shared__ double red[4][TPB / 32];
      const grid_t &g = P.g;
      double pud[4] = {0, 0, 0, 0};   // liquid volume, dry volume, liquid number, particle number leaving through z0

      // grid-stride over tiles of TPB super-droplets: the precipitation sums are reduced once per CTA
#ifndef LCX_TR_NO_PREFETCH
      // the cell index and the permutation entry of the NEXT tile are asked for at the top of this one: the two dependent load levels
      // of a super-droplet (index -> Courant numbers / position) then start one tile apart instead of back to back (1.83 -> 1.77 ms;
      // also asking the next tile's permuted positions into L2 at the end of the body: 1.79, dropped)
      const size_t t0 = size_t(blockIdx.x) * TPB + threadIdx.x, stride = size_t(gridDim.x) * TPB;
      idx_t c_next = t0 < n_part ? ijk[t0] : 0;
      uint32_t src_next = (LAZY && t0 < n_part) ? perm[t0] : 0;
      for (size_t t = t0; t < n_part; t += stride)
      {
        const idx_t c = c_next;
        const uint32_t src_now = src_next;
        if (t + stride < n_part) { c_next = ijk[t + stride]; if (LAZY) src_next = perm[t + stride]; }
#else
      for (size_t t = size_t(blockIdx.x) * TPB + threadIdx.x; t < n_part; t += size_t(gridDim.x) * TPB)
      {
        const idx_t c = ijk[t];
#endif
        idx_t i = 0, j = 0, k = 0;
        switch (g.n_dims)      // two integer divisions at most: q = c / nz, then i = q / ny
        {
          case 1: i = c; break;
          case 2: i = fastdiv(c, P.nz_m, P.nz_s1, P.nz_s2); k = c - i * g.nz; break;
          case 3: { const idx_t q = fastdiv(c, P.nz_m, P.nz_s1, P.nz_s2); k = c - q * g.nz; i = fastdiv(q, P.ny_m, P.ny_s1, P.ny_s2); j = q - i * g.ny; } break;
        }
        real_t x, y, z;
#ifndef LCX_TR_NO_PREFETCH
        if (LAZY) { const uint32_t src = src_now; x = g.nx ? xi[src] : 0; y = g.ny ? yi[src] : 0; z = g.nz ? zi[src] : 0; }
#else
        if (LAZY) { const uint32_t src = perm[t]; x = g.nx ? xi[src] : 0; y = g.ny ? yi[src] : 0; z = g.nz ? zi[src] : 0; }
#endif
        else      { x = g.nx ? xs[t] : 0; y = g.ny ? ys[t] : 0; z = g.nz ? zs[t] : 0; }
        n_t n = ns[t];
        const n_t n_in = n;

        if (P.adve && g.n_dims > 0)
        {
          if (P.scheme == AS_IMPLICIT || P.scheme == AS_EULER)
          {
            const faces f = faces_of(g, c + g.halo_x, i + idx_t(g.halo_size), j);
            if (P.scheme == AS_IMPLICIT)
            {
              x = step_impl(x, i, Cx[f.xl], Cx[f.xr], g.dx);
              if (g.n_dims > 2) y = step_impl(y, j, Cy[f.yl], Cy[f.yr], g.dy);
              if (g.n_dims > 1) z = step_impl(z, k, Cz[f.zl], Cz[f.zr], g.dz);
            }
            else
            {
              x = step_expl(x, i, Cx[f.xl], Cx[f.xr], g.dx, true);
              if (g.n_dims > 2) y = step_expl(y, j, Cy[f.yl], Cy[f.yr], g.dy, true);
              if (g.n_dims > 1) z = step_expl(z, k, Cz[f.zl], Cz[f.zr], g.dz, true);
            }
          }
          else   // predictor-corrector in halo-shifted coordinates: adve.ipp:183-303
          {
            x = x + real_t(g.halo_size) * g.dx;
            idx_t ih, jh, kh;
            idx_t ch = cell_of(P, x, y, z, ih, jh, kh);
            real_t x_old = x, y_old = y, z_old = z;
            faces f = faces_of(g, ch, ih, jh);
            x = step_expl(x, ih, Cx[f.xl], Cx[f.xr], g.dx, true);
            if (g.n_dims > 2) y = step_expl(y, jh, Cy[f.yl], Cy[f.yr], g.dy, true);
            if (g.n_dims > 1) z = step_expl(z, kh, Cz[f.zl], Cz[f.zr], g.dz, true);
            if (g.n_dims > 1)
            {
              if (z >= g.z1) z = g.z1 - 1e-8 * g.dz;
              if (z <= g.z0) z = g.z0 + 1e-8 * g.dz;
            }
            if (g.n_dims == 3)
            {
              if (y >= g.y1) y_old = y_old + (g.y1 - g.y0);
              if (y < g.y0)  y_old = y_old - (g.y1 - g.y0);
              y = periodic_wrap(y, g.y0, g.y1, P.inv_Ly);
            }
            ch = cell_of(P, x, y, z, ih, jh, kh);
            x_old = x + x_old;
            if (g.n_dims > 2) y_old = y + y_old;
            if (g.n_dims > 1) z_old = z + z_old;
            f = faces_of(g, ch, ih, jh);
            x = step_expl(x, ih, Cx[f.xl], Cx[f.xr], g.dx, false);
            if (g.n_dims > 2) y = step_expl(y, jh, Cy[f.yl], Cy[f.yr], g.dy, false);
            if (g.n_dims > 1) z = step_expl(z, kh, Cz[f.zl], Cz[f.zr], g.dz, false);
            x = (x + x_old) / real_t(2.);
            if (g.n_dims > 2) y = (y + y_old) / real_t(2.);
            if (g.n_dims > 1) z = (z + z_old) / real_t(2.);
            x = x - real_t(g.halo_size) * g.dx;
          }
        }

        if (P.sedi) z = z - P.dt * vt[t];            // sedi.ipp:18-23 (vt may be the -1 "invalid" marker, as in the reference)
        if (P.subs) z = z - P.dt * w_LS[k];          // subs.ipp:18-23 (k from before the move)

        uint32_t fl = 0;
        if (g.n_dims > 0)
        {
          // x walls: bcnd.ipp:124-195
          if (P.bcond_lft == LCX_BCOND_SHAREDMEM && P.bcond_rgt == LCX_BCOND_SHAREDMEM)
          {
            if (!P.open_side_walls) x = periodic_wrap(x, g.x0, g.x1, P.inv_Lx);
            else if (x >= g.x1 || x < g.x0) n = 0;
          }
          else
          {
            if (x < g.x0)       { if (P.bcond_lft == LCX_BCOND_OPEN) n = 0; else fl = 1; }
            else if (x >= g.x1) { if (P.bcond_rgt == LCX_BCOND_OPEN) n = 0; else fl = 2; }
          }
          // y walls: bcnd.ipp:197-219
          if (g.n_dims == 3)
          {
            if (!P.open_side_walls) y = periodic_wrap(y, g.y0, g.y1, P.inv_Ly);
            else if (y >= g.y1 || y < g.y0) n = 0;
          }
          // z walls: bcnd.ipp:221-364
          if (g.n_dims > 1)
          {
            if (!P.periodic_topbot)
            {
              if (z >= g.z1)
              {
                // leaves through the lid without any accounting in the reference (bcnd.ipp:330-336); freshly collided SDs do so
                // regularly: their fall speed is the "invalid" marker -1, i.e. they rise by dt metres.  Rare, so the tally for
                // conservation checks (lcx_top_loss) is a plain atomic - it feeds no result.
                if (n != 0) { atomicAdd(M.top_loss, double(count_vol(real_t(n), rd3[t], real_t(1.)))); atomicAdd(M.top_loss + 1, 1.0); }
                n = 0;
              }
              if (z < g.z0)
              {
                const real_t nf = real_t(n);
                const real_t r2 = rw2[t];
                pud[0] += count_vol(nf, r2, real_t(3. / 2.));
                pud[1] += count_vol(nf, rd3[t], real_t(1.));
                pud[2] += (r2 == real_t(0)) ? 0. : nf;
                pud[3] += nf;
                n = 0;
              }
            }
            else z = periodic_wrap(z, g.z0, g.z1, P.inv_Lz);
          }
        }

        if (g.nx) xs[t] = x;
        if (g.ny) ys[t] = y;
        if (g.nz) zs[t] = z;
        if (n != n_in) ns[t] = n;      // only SDs that left the domain change their multiplicity here
        if (fl)
        {
          const unsigned slot = atomicAdd(M.count[fl - 1], 1u);
          if (slot < M.cap) { M.key[fl - 1][slot] = M.sid[t]; M.val[fl - 1][slot] = uint32_t(t); }     // overflow is reported by lcx_migr_put
        }
        // sort key of the coming re-layout (hskpng_ijk of post_copy): new cell, or n_cell for SDs that are gone;
        // migrants leave through lcx_migr_pack, which zeroes their multiplicity and re-keys them
        uint32_t kx = relayout_dead_key(g);
        if (n != 0 && fl == 0)
        {
          idx_t i2, j2, k2;
          idx_t cell = cell_of(P, x, y, z, i2, j2, k2);
          if (cell >= g.n_cell) cell = g.n_cell - 1;
          kx = relayout_key(g, cell, g.class_bits ? rw2[t] : real_t(0));
        }
        key[t] = kx;      // the identity permutation that goes with the keys is only written when the full sort needs it (lcx_layout.cu)
      }

      // deterministic block sums of the precipitation terms (second pass: k_puddle_final)
#pragma unroll
      for (int q = 0; q < 4; ++q)
      {
        double v = pud[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) red[q][threadIdx.x >> 5] = v;
      }
      __syncthreads();
      if (threadIdx.x < 4)
      {
        double s = 0;
        for (int w = 0; w < TPB / 32; ++w) s += red[threadIdx.x][w];
        partial[size_t(threadIdx.x) * gridDim.x + blockIdx.x] = s;
      }
    }

    __global__ void __launch_bounds__(TPB) k_puddle_final(unsigned n_blocks, const double *__restrict__ partial, dev_scalars *sc)
    {
      __shared__ double red[TPB];
      for (int q = 0; q < 4; ++q)
      {
        double s = 0;
        for (unsigned b = threadIdx.x; b < n_blocks; b += TPB) s += partial[size_t(q) * n_blocks + b];
        red[threadIdx.x] = s;
        __syncthreads();
        for (int o = TPB / 2; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
        if (threadIdx.x == 0) sc->puddle[q] += red[0];
        __syncthreads();
      }
    }

    // ---- migration ---------------------------------------------------------------------------------------
    struct mig_attrs { const real_t *src[12]; int n; int x_slot; };

    // Packs the leavers of one side STRAIGHT into the neighbour's inbox (out_n / out_real may be peer memory: plain coalesced
    // 8-byte stores over NVLink), attribute-major like the reference's buffers (pack.ipp:37-46,77-86).
    __global__ void __launch_bounds__(TPB) k_mig_pack(unsigned count, const uint32_t *__restrict__ val, mig_attrs A, n_t *__restrict__ ns,
                                                     real_t lcl, real_t rmt, n_t *__restrict__ out_n, real_t *__restrict__ out_real)
    {
      const unsigned jx = blockIdx.x * TPB + threadIdx.x;
      if (jx >= count) return;
      const uint32_t ph = val[jx];
      out_n[jx] = ns[ph];
      for (int a = 0; a < A.n; ++a)
      {
        real_t v = A.src[a][ph];
        if (a == A.x_slot) v = rmt + v - lcl;                 // remote coordinate: pack.ipp:15-26
        out_real[size_t(a) * count + jx] = v;
      }
      ns[ph] = 0;                                             // flag_lft / flag_rgt: unpack.ipp:122-145 (its sort key already says "gone")
    }

    // publishes a delivery: the data written by the kernels before this one on the same stream are made visible system-wide
    // first, then {count, seq} appear in the receiver's header
    __global__ void k_mig_signal(mig_hdr *hdr, unsigned count, unsigned seq)
    {
      hdr->count = count;
      __threadfence_system();
      *reinterpret_cast<volatile unsigned int *>(&hdr->seq) = seq;
      __threadfence_system();
    }

    // waits (on the device) until the deliveries with sequence number `seq` have been published in this engine's own inbox
    // headers; only used when the neighbour lives in another process (other GPU: its progress does not depend on this stream).
    // Gives up after ~20 s instead of hanging the GPU.
    __global__ void k_mig_wait(const mig_hdr *h0, const mig_hdr *h1, unsigned seq, dev_scalars *sc, long long max_cycles)
    {
      const long long t0 = clock64();
      for (int q = 0; q < 2; ++q)
      {
        const mig_hdr *h = q == 0 ? h0 : h1;
        if (!h) continue;
        while (*reinterpret_cast<const volatile unsigned int *>(&h->seq) != seq)
        {
          __nanosleep(200);
          if (clock64() - t0 > max_cycles) { sc->mig_timeout = 1u; return; }
        }
      }
      __threadfence_system();
    }

    // ---- Courant halo planes between process-distributed slabs: particles_impl_xchng_courants.ipp:15-153 -----------------
    // up to three contiguous pieces (Cx, Cy, Cz planes) copied between this engine's Courant arrays and a packed buffer that
    // may be peer memory (the neighbour's inbox): plain coalesced stores over NVLink
    struct halo_pieces { const real_t *src[3]; real_t *dst[3]; unsigned n[3]; };
    __global__ void __launch_bounds__(TPB) k_halo_copy(halo_pieces P)
    {
      unsigned t = blockIdx.x * TPB + threadIdx.x;
      for (int q = 0; q < 3; ++q)
      {
        if (t < P.n[q]) { P.dst[q][t] = P.src[q][t]; return; }
        t -= P.n[q];
      }
    }
    __global__ void k_halo_signal(mig_hdr *hdr, unsigned seq)
    {
      __threadfence_system();
      *reinterpret_cast<volatile unsigned int *>(&hdr->halo_seq) = seq;
      __threadfence_system();
    }
    __global__ void k_halo_wait(const mig_hdr *h0, const mig_hdr *h1, unsigned seq, dev_scalars *sc, long long max_cycles)
    {
      const long long t0 = clock64();
      for (int q = 0; q < 2; ++q)
      {
        const mig_hdr *h = q == 0 ? h0 : h1;
        if (!h) continue;
        while (*reinterpret_cast<const volatile unsigned int *>(&h->halo_seq) != seq)
        {
          __nanosleep(200);
          if (clock64() - t0 > max_cycles) { sc->mig_timeout = 2u; return; }
        }
      }
      __threadfence_system();
    }

    struct mig_dst { real_t *dst[12]; int n; int x_slot; };

    __global__ void __launch_bounds__(TPB) k_mig_unpack(unsigned count, size_t n_part_old, size_t sid_first, mig_dst A, n_t *__restrict__ ns, idx_t *__restrict__ sid,
                                                       const n_t *__restrict__ in_n, const real_t *__restrict__ in_real,
                                                       real_t x0, real_t x1, real_t tol, int from_right)
    {
      const unsigned jx = blockIdx.x * TPB + threadIdx.x;
      if (jx >= count) return;
      const size_t d = n_part_old + jx;
      ns[d] = in_n[jx];
      sid[d] = idx_t(sid_first + jx);
      for (int a = 0; a < A.n; ++a)
      {
        real_t v = in_real[size_t(a) * count + jx];
        if (a == A.x_slot)
        {
          v = v >= x1 ? v - tol : v < x0 ? v + tol : v;       // tolerance_away_from_bcond: unpack.ipp:15-31,101
          if (from_right && v == x1) v = nextafter(v, real_t(0.));   // step_async_and_copy.ipp:137
        }
        A.dst[a][d] = v;
      }
    }

    int fill_attr_list(lcx_engine *e, real_t **list, int *x_slot)
    {
      sd_arrays &s = e->S();
      int n = 0;
      list[n++] = s.rd3.p; list[n++] = s.rw2.p; list[n++] = s.kpa.p; list[n++] = s.vt.p;
      *x_slot = -1;
      if (e->grid.nx) { *x_slot = n; list[n++] = s.x.p; }
      if (e->grid.ny) list[n++] = s.y.p;
      if (e->grid.nz) list[n++] = s.z.p;
      if (s.pp_rv.p) { list[n++] = s.pp_rv.p; list[n++] = s.pp_th.p; list[n++] = s.pp_rh.p; if (s.pp_p.p) list[n++] = s.pp_p.p; }   // particles_impl.ipp:452-459
      if (s.rc2.p) list[n++] = s.rc2.p;                                                                                           // particles_impl.ipp:488-491
      return n;
    }
  }

  void transport(lcx_engine *e, const lcx_transport_opts *o)
  {
    const size_t n = e->n_part;
    if (n == 0 || e->grid.n_dims == 0) return;
    wait_courant(e);
    sd_arrays &s = e->S();
    tr_params P;
    P.g = e->grid;
    P.adve = o->adve; P.sedi = o->sedi && e->grid.nz; P.subs = o->subs && e->grid.nz; P.scheme = o->adve_scheme;
    P.dt = real_t(o->dt);
    P.open_side_walls = e->cfg.open_side_walls; P.periodic_topbot = e->cfg.periodic_topbot_walls;
    P.bcond_lft = e->cfg.bcond_lft; P.bcond_rgt = e->cfg.bcond_rgt;
    {
      const grid_t &g = e->grid;
      P.inv_dx = g.nx ? real_t(1) / g.dx : real_t(0); P.inv_dy = g.ny ? real_t(1) / g.dy : real_t(0); P.inv_dz = g.nz ? real_t(1) / g.dz : real_t(0);
      P.inv_Lx = g.nx ? real_t(1) / (g.x1 - g.x0) : real_t(0); P.inv_Ly = g.ny ? real_t(1) / (g.y1 - g.y0) : real_t(0); P.inv_Lz = g.nz ? real_t(1) / (g.z1 - g.z0) : real_t(0);
      fastdiv_setup(uint32_t(g.nz > 0 ? g.nz : 1), P.nz_m, P.nz_s1, P.nz_s2);
      fastdiv_setup(uint32_t(g.ny > 0 ? g.ny : 1), P.ny_m, P.ny_s1, P.ny_s2);
    }
    if (P.subs && e->w_LS.n < size_t(e->grid.nz)) throw error("subsidence requested but no w_LS profile was set");
    if (P.scheme == AS_PRED_CORR && e->grid.halo_size != 2) throw error("predictor-corrector advection needs a 2-cell Courant halo");
    static const int ctas_per_sm = [] { const char *v = std::getenv("LCX_TR_CTAS"); return v ? std::atoi(v) : 48; }();
    const unsigned blocks = unsigned(ctas_per_sm > 0 ? std::min<size_t>(div_up(n, TPB), size_t(148) * ctas_per_sm) : div_up(n, TPB));
    if (e->red_partial.n < size_t(blocks) * 4) { LCX_CUDA(cudaStreamSynchronize(e->stream)); e->red_partial.alloc(size_t(blocks) * 4 + 1024); }
    mig_lists M = {};
    M.sid = s.sid.p; M.cap = unsigned(e->mig_cap);
    M.count[0] = &e->scalars.p->n_lft; M.count[1] = &e->scalars.p->n_rgt; M.top_loss = &e->scalars.p->puddle[4];
    for (int sd = 0; sd < 2; ++sd) { M.key[sd] = e->mig_key[sd][0].p; M.val[sd] = e->mig_val[sd][0].p; }
    if (P.bcond_lft == LCX_BCOND_DISTMEM || P.bcond_rgt == LCX_BCOND_DISTMEM)
      LCX_CUDA(cudaMemsetAsync(&e->scalars.p->n_lft, 0, 2 * sizeof(unsigned int), e->stream));
    if (e->pending & lcx_engine::PENDING_XYZ)
    {
      sd_arrays &o2 = e->A();
      e->pending &= ~unsigned(lcx_engine::PENDING_XYZ);
      LCX_LAUNCH(e, k_transport<true>, blocks, TPB, 0, n, P, s.x.p, s.y.p, s.z.p, s.vt.p, s.ijk.p, s.n.p, s.rw2.p, s.rd3.p,
                 e->courant_x.p, e->courant_y.p, e->courant_z.p, e->w_LS.p, e->red_partial.p, e->key[0].p,
                 e->pending_perm.p, o2.x.p, o2.y.p, o2.z.p, M);
    }
    else
      LCX_LAUNCH(e, k_transport<false>, blocks, TPB, 0, n, P, s.x.p, s.y.p, s.z.p, s.vt.p, s.ijk.p, s.n.p, s.rw2.p, s.rd3.p,
                 e->courant_x.p, e->courant_y.p, e->courant_z.p, e->w_LS.p, e->red_partial.p, e->key[0].p,
                 nullptr, nullptr, nullptr, nullptr, M);
    e->keys_ready = n;      // key[0] / val[0] hold the sort keys of SDs [0, n)
    if (e->grid.n_dims > 1 && !e->cfg.periodic_topbot_walls)
      LCX_LAUNCH(e, k_puddle_final, 1, TPB, 0, blocks, e->red_partial.p, e->scalars.p);
    e->grouped = false;   // positions changed: cell segments are stale until lcx_post_copy
  }

  // Courant values of the planes one neighbour needs: halo_size x-planes of each staggered field
  static void halo_counts(const grid_t &g, unsigned n[3])
  {
    const unsigned h = unsigned(g.halo_size);
    n[0] = n[1] = n[2] = 0;
    if (h == 0 || g.n_dims == 0) return;
    n[0] = g.n_dims == 1 ? h : g.n_dims == 2 ? h * g.nz : h * g.nz * g.ny;                       // halo_x
    if (g.n_dims == 3) n[1] = h * (g.ny + 1) * g.nz;                                            // halo_y
    if (g.n_dims >= 2) n[2] = g.n_dims == 2 ? h * (g.nz + 1) : h * (g.nz + 1) * g.ny;           // halo_z
  }
  size_t halo_values(const grid_t &g) { unsigned n[3]; halo_counts(g, n); return size_t(n[0]) + n[1] + n[2]; }

  static long long wait_cycles()
  {
    static const long long c = [] { const char *v = std::getenv("LCX_MIG_TIMEOUT_S"); return (long long)((v ? std::atof(v) : 20.) * 2e9); }();
    return c;
  }

  // Each distributed side: this slab's outermost interior planes go STRAIGHT into the neighbour's inbox (its halo section), then
  // the delivery is published.  What goes where is the reference's rule (xchng_courants.ipp:26-52): to the left neighbour the
  // planes right of face 0 (Cx: faces 1..h, Cy / Cz: columns 0..h-1) - its right halo; to the right neighbour the last h planes
  // (Cx: faces nx-h..nx-1, Cy / Cz: columns nx-h..nx-1) - its left halo.
  void halo_put(lcx_engine *e)
  {
    const grid_t &g = e->grid;
    unsigned n[3];
    halo_counts(g, n);
    const size_t total = size_t(n[0]) + n[1] + n[2];
    const bool dm[2] = {e->cfg.bcond_lft == LCX_BCOND_DISTMEM, e->cfg.bcond_rgt == LCX_BCOND_DISTMEM};
    if (total == 0 || (!dm[0] && !dm[1])) return;
    wait_courant(e);
    const unsigned seq = ++e->halo_seq;
    const int parity = int(seq & 1u);
    const size_t h = size_t(g.halo_size);
    const size_t col_x = n[0] / h, col_y = n[1] / h, col_z = n[2] / h;      // values per x-plane
    for (int side = 0; side < 2; ++side)
    {
      if (!dm[side]) continue;
      lcx_engine::mig_remote &r = e->remote[side];
      if (!r.base) throw error("lcx_halo_put: no neighbour connected on side " + std::to_string(side));
      real_t *out = box_halo(r.base, r.cap, e->mig_n_real, total, parity);
      halo_pieces P;
      // first plane sent, in planes of the halo-extended arrays (interior face / column f sits at plane h + f)
      const size_t px = side == 0 ? h + 1 : size_t(g.nx), pyz = side == 0 ? h : size_t(g.nx);
      P.src[0] = e->courant_x.p + px * col_x;                           P.dst[0] = out;               P.n[0] = n[0];
      P.src[1] = n[1] ? e->courant_y.p + pyz * col_y : nullptr;         P.dst[1] = out + n[0];        P.n[1] = n[1];
      P.src[2] = n[2] ? e->courant_z.p + pyz * col_z : nullptr;         P.dst[2] = out + n[0] + n[1]; P.n[2] = n[2];
      LCX_LAUNCH(e, k_halo_copy, div_up(total, TPB), TPB, 0, P);
      LCX_LAUNCH(e, k_halo_signal, 1, 1, 0, box_hdr(r.base, parity), seq);
    }
  }

  // Waits (on the device) for both neighbours' planes and copies them into this slab's halo: the right neighbour's into the last
  // h planes of each array, the left neighbour's into the first h
  void halo_take(lcx_engine *e)
  {
    const grid_t &g = e->grid;
    unsigned n[3];
    halo_counts(g, n);
    const size_t total = size_t(n[0]) + n[1] + n[2];
    // inbox 0 is filled by the right neighbour, inbox 1 by the left one
    const bool dm[2] = {e->cfg.bcond_rgt == LCX_BCOND_DISTMEM, e->cfg.bcond_lft == LCX_BCOND_DISTMEM};
    if (total == 0 || (!dm[0] && !dm[1])) return;
    const unsigned seq = e->halo_seq;
    const int parity = int(seq & 1u);
    LCX_LAUNCH(e, k_halo_wait, 1, 1, 0, dm[0] ? box_hdr(e->inbox[0].p, parity) : nullptr, dm[1] ? box_hdr(e->inbox[1].p, parity) : nullptr,
               seq, e->scalars.p, wait_cycles());
    for (int q = 0; q < 2; ++q)
    {
      if (!dm[q]) continue;
      const real_t *in = box_halo(e->inbox[q].p, e->mig_cap, e->mig_n_real, total, parity);
      halo_pieces P;
      P.src[0] = in;               P.dst[0] = q == 0 ? e->courant_x.p + e->courant_x.n - n[0] : e->courant_x.p;                 P.n[0] = n[0];
      P.src[1] = in + n[0];        P.dst[1] = n[1] ? (q == 0 ? e->courant_y.p + e->courant_y.n - n[1] : e->courant_y.p) : nullptr; P.n[1] = n[1];
      P.src[2] = in + n[0] + n[1]; P.dst[2] = n[2] ? (q == 0 ? e->courant_z.p + e->courant_z.n - n[2] : e->courant_z.p) : nullptr; P.n[2] = n[2];
      LCX_LAUNCH(e, k_halo_copy, div_up(total, TPB), TPB, 0, P);
    }
  }

  // Sorts each side's leavers by storage index (the reference lists them in ascending SD index: bcnd.ipp:160-172), packs them
  // into the neighbour's inbox and publishes the delivery.  One host read-back: the two counts.
  static bool mig_debug() { static const bool on = [] { const char *v = std::getenv("LCX_MIG_DEBUG"); return v && v[0] == '1'; }(); return on; }

  void migr_put(lcx_engine *e, int64_t *n_lft, int64_t *n_rgt)
  {
    sd_arrays &s = e->S();
    *n_lft = *n_rgt = 0;
    const bool dm[2] = {e->cfg.bcond_lft == LCX_BCOND_DISTMEM, e->cfg.bcond_rgt == LCX_BCOND_DISTMEM};
    if (!dm[0] && !dm[1]) return;
    const unsigned seq = ++e->mig_seq;
    const int parity = int(seq & 1u);
    unsigned count[2] = {0, 0};
    if (e->n_part)
    {
      LCX_CUDA(cudaMemcpyAsync(e->h_scalars, e->scalars.p, sizeof(dev_scalars), cudaMemcpyDeviceToHost, e->stream));
      LCX_CUDA(cudaStreamSynchronize(e->stream));
      count[0] = e->h_scalars->n_lft; count[1] = e->h_scalars->n_rgt;
    }
    for (int side = 0; side < 2; ++side)
    {
      if (!dm[side]) { count[side] = 0; continue; }
      lcx_engine::mig_remote &r = e->remote[side];
      if (!r.base) throw error("lcx_migr_put: no neighbour connected on side " + std::to_string(side) + " (lcx_migr_connect / lcx_migr_ipc_connect)");
      const unsigned cnt = count[side];
      if (cnt > e->mig_cap || cnt > r.cap)
        throw error("migration buffer overflow: " + std::to_string(cnt) + " super-droplets cross one slab face, capacity " + std::to_string(e->mig_cap < r.cap ? e->mig_cap : r.cap));
      if (cnt)
      {
        uint32_t *mk[2] = {e->mig_key[side][0].p, e->mig_key[side][1].p}, *mv[2] = {e->mig_val[side][0].p, e->mig_val[side][1].p};
        int bits = 0; { uint64_t v = e->sid_hi ? e->sid_hi - 1 : 0; while (v) { ++bits; v >>= 1; } if (bits == 0) bits = 1; }
        const int res = radix_sort_pairs(e, cnt, 0, bits, mk, mv, 0);
        mig_attrs A; real_t *list[12];
        A.n = fill_attr_list(e, list, &A.x_slot);
        for (int a = 0; a < A.n; ++a) A.src[a] = list[a];
        const real_t lcl = side == 0 ? e->grid.x0 : e->grid.x1;
        const real_t rmt = side == 0 ? real_t(e->cfg.lft_x1) : real_t(e->cfg.rgt_x0);
        LCX_LAUNCH(e, k_mig_pack, div_up(cnt, TPB), TPB, 0, cnt, mv[res], A, s.n.p, lcl, rmt,
                   box_n(r.base, r.cap, parity), box_real(r.base, r.cap, e->mig_n_real, parity));
      }
      LCX_LAUNCH(e, k_mig_signal, 1, 1, 0, box_hdr(r.base, parity), cnt, seq);
    }
    LCX_CUDA(cudaEventRecord(e->ev_put, e->stream));
    *n_lft = count[0]; *n_rgt = count[1];
    if (mig_debug()) std::fprintf(stderr, "[mig dev %d] put seq %u: %u left, %u right (n_part %zu)\n", e->device, seq, count[0], count[1], e->n_part);
  }

  // Appends the arrivals: first the right neighbour's left-movers, then the left neighbour's right-movers, each in the
  // sender's storage order (step_async_and_copy.ipp:100-190).  One host read-back: the two counts.
  void migr_take(lcx_engine *e, lcx_engine *rgt, lcx_engine *lft, int64_t *n_from_rgt, int64_t *n_from_lft)
  {
    *n_from_rgt = *n_from_lft = 0;
    // inbox 0 is filled by the right neighbour, inbox 1 by the left one
    const bool dm[2] = {e->cfg.bcond_rgt == LCX_BCOND_DISTMEM, e->cfg.bcond_lft == LCX_BCOND_DISTMEM};
    if (!dm[0] && !dm[1]) return;
    const unsigned seq = e->mig_seq;
    const int parity = int(seq & 1u);
    lcx_engine *nb[2] = {rgt, lft};
    const mig_hdr *wait_for[2] = {nullptr, nullptr};
    for (int q = 0; q < 2; ++q)
    {
      if (!dm[q]) continue;
      if (nb[q]) { if (nb[q] != e) LCX_CUDA(cudaStreamWaitEvent(e->stream, nb[q]->ev_put, 0)); }
      else wait_for[q] = box_hdr(e->inbox[q].p, parity);
    }
    const long long max_cycles = wait_cycles();
    if (wait_for[0] || wait_for[1])
      LCX_LAUNCH(e, k_mig_wait, 1, 1, 0, wait_for[0], wait_for[1], seq, e->scalars.p, max_cycles);
    mig_hdr h[2] = {};
    for (int q = 0; q < 2; ++q)
      if (dm[q]) LCX_CUDA(cudaMemcpyAsync(&h[q], box_hdr(e->inbox[q].p, parity), sizeof(mig_hdr), cudaMemcpyDeviceToHost, e->stream));
    LCX_CUDA(cudaMemcpyAsync(e->h_scalars, e->scalars.p, sizeof(dev_scalars), cudaMemcpyDeviceToHost, e->stream));
    LCX_CUDA(cudaStreamSynchronize(e->stream));
    if (mig_debug()) std::fprintf(stderr, "[mig dev %d] take seq %u: headers (%u, %u) (%u, %u) timeout %u\n", e->device, seq, h[0].seq, h[0].count, h[1].seq, h[1].count, e->h_scalars->mig_timeout);
    if (e->h_scalars->mig_timeout == 2u)
      throw error("Courant halo exchange: a neighbour's planes did not arrive in time (predictor-corrector advection between process-distributed slabs; $LCX_MIG_TIMEOUT_S)");
    if (e->h_scalars->mig_timeout)
      throw error("x-slab migration: a neighbour's delivery did not arrive in time (waiting for sequence number " + std::to_string(seq) + "; inbox from the right holds " +
                  std::to_string(h[0].seq) + " with " + std::to_string(h[0].count) + " super-droplets, inbox from the left " + std::to_string(h[1].seq) + " with " +
                  std::to_string(h[1].count) + "; $LCX_MIG_TIMEOUT_S)");
    sd_arrays &s = e->S();
    for (int q = 0; q < 2; ++q)
    {
      if (!dm[q]) continue;
      if (h[q].seq != seq) throw error("x-slab migration: sequence mismatch (neighbours out of step: got " + std::to_string(h[q].seq) + ", expected " + std::to_string(seq) + ")");
      const size_t count = h[q].count;
      (q == 0 ? *n_from_rgt : *n_from_lft) = int64_t(count);
      if (count == 0) continue;
      if (count > e->mig_cap) throw error("migration buffer overflow on receive");
      if (e->n_part + count > e->cap)
        throw error("n_sd_max (" + std::to_string(e->cap) + ") < n_part (" + std::to_string(e->n_part + count) + ")");
      mig_dst A; real_t *list[12];
      A.n = fill_attr_list(e, list, &A.x_slot);
      for (int a = 0; a < A.n; ++a) A.dst[a] = list[a];
      if (e->sid_hi + count > e->cap) densify_sid(e);
      LCX_LAUNCH(e, k_mig_unpack, div_up(count, TPB), TPB, 0, unsigned(count), e->n_part, e->sid_hi, A, s.n.p, s.sid.p,
                 box_n(e->inbox[q].p, e->mig_cap, parity), box_real(e->inbox[q].p, e->mig_cap, e->mig_n_real, parity),
                 e->grid.x0, e->grid.x1, real_t(5e-4), int(q == 0));
      e->n_part += count;
      e->sid_hi += count;
      e->grouped = false;
    }
  }
}
