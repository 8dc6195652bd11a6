// C++ host layer of the B200-native Lagrangian microphysics: the `lgrngn::factory` / `particles_proto_t`
// implementation for the CUDA and multi_CUDA back-ends.  All device work goes through the C ABI of
// include/lcx_b200.h (liblcx_b200.so); nothing here touches CUDA directly.
//
// What lives here (and which reference code it stands in for):
//   * option checks, call-order state machine, error texts   src/particles_step.ipp:32-494, src/particles_init.ipp:16-131,
//                                                            src/impl/initialization/particles_impl_init_sanity_check.ipp:14-176
//   * Eulerian <-> Lagrangian index maps and field syncing    src/impl/initialization/particles_impl_init_e2l.ipp:34-114,
//                                                            src/impl/particles_impl_sync.ipp:15-68
//   * super-droplet initialisation from dry spectra            src/impl/initialization/particles_impl_init_{dist_analysis,
//     (host side, same libm and the same mt19937 draw order     SD_with_distros_sd_conc,count_num,ijk,dry_sd_conc,n,wet,xyz}.ipp
//     as the reference's CPU back-ends => bit-identical state)
//   * random streams for coalescence                           src/detail/urand.hpp:19-88 (replay) or Philox in the kernels
//   * x-slab decomposition and migration between devices        src/detail/distmem_opts.hpp:10-52,
//                                                            src/impl_multi_gpu/particles_multi_gpu_impl.ipp:15-207,
//                                                            src/impl_multi_gpu/particles_multi_gpu_impl_step_async_and_copy.ipp:28-206
#include <libcloudph++/lgrngn/factory.hpp>

#include "lcx_b200.h"
#include "lcx_api_select.hpp"
#include "lcx_physics.h"
#include "particles_b200.h"
#include <lgrngn_abi_probe.hpp>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <dlfcn.h>
#include <exception>
#include <functional>
#include <mutex>
#include <thread>
#include <fstream>
#include <iostream>
#include <limits>
#include <random>
#include <sstream>

namespace libcloudphxx
{
  namespace lgrngn
  {
    namespace b200
    {
      // ---- process-wide settings (see particles_b200.h) ---------------------------------------------------
      namespace
      {
        int initial_rng_mode()
        {
          const char *s = std::getenv("LCX_RNG");
          if (s && (std::string(s) == "mt19937" || std::string(s) == "replay")) return LGRNGN_B200_RNG_MT19937;
          return LGRNGN_B200_RNG_PHILOX;
        }
        int g_rng_mode = initial_rng_mode();
        int g_dense_sid = -1;          // -1: dense storage indices exactly when the random stream is injected; 0 / 1: forced
        lgrngn_b200_distmem g_distmem = {0, 1, -1., -1., 0};

        void chk(int rc) { if (rc != 0) throw std::runtime_error(std::string("libcloudph++ (B200 engine): ") + lcx_last_error()); }

        std::string data_dir()
        {
          if (const char *s = std::getenv("LCX_DATA_DIR")) return s;
          Dl_info info;
          if (dladdr(reinterpret_cast<void *>(&initial_rng_mode), &info) && info.dli_fname)
          {
            std::string p(info.dli_fname);
            const size_t slash = p.find_last_of('/');
            p = (slash == std::string::npos) ? std::string(".") : p.substr(0, slash);
            return p + "/../data";
          }
          return "data";
        }

        template <class real_t>
        std::vector<real_t> load_efficiencies(const std::string &name)
        {
          const std::string path = data_dir() + "/" + name + ".f64";
          std::ifstream f(path, std::ios::binary);
          if (!f) throw std::runtime_error("libcloudph++: cannot open collision-efficiency table " + path);
          f.seekg(0, std::ios::end);
          const size_t bytes = size_t(f.tellg());
          f.seekg(0);
          std::vector<double> raw(bytes / sizeof(double));
          f.read(reinterpret_cast<char *>(raw.data()), std::streamsize(raw.size() * sizeof(double)));
          return std::vector<real_t>(raw.begin(), raw.end());
        }

        inline int m1(int n) { return n == 0 ? 1 : n; }
      }

      enum bcond_t { sharedmem = LCX_BCOND_SHAREDMEM, distmem = LCX_BCOND_DISTMEM, open = LCX_BCOND_OPEN };

      // x-slab decomposition: src/detail/distmem_opts.hpp:10-52
      template <class real_t>
      int get_dev_nx(const opts_init_t<real_t> &o, int rank, int size)
      {
        if (rank < size - 1) return int(o.nx / size + .5);
        return o.nx - rank * int(o.nx / size + .5);
      }
      template <class real_t>
      int distmem_opts(opts_init_t<real_t> &o, int rank, int size)
      {
        const int n_x_bfr = rank * get_dev_nx(o, 0, size);
        o.nx = get_dev_nx(o, rank, size);
        if (rank != 0) o.x0 = 0.;
        if (rank != size - 1) o.x1 = o.nx * o.dx;
        else o.x1 = o.x1 - n_x_bfr * o.dx;
        o.n_sd_max = o.n_sd_max / size + 1;
        return n_x_bfr;
      }

      // One host thread per slab, kept for the life time of the particle system.  The reference's multi_CUDA spawns a
      // std::thread per device in every API call (particles_multi_gpu_impl.ipp:208-227); here the threads persist and each
      // call hands them one job.  Every slab's blocking points (read-backs of counts, field copies) then only stall its own
      // thread, so the GPUs run concurrently.
      class slab_workers
      {
        struct worker
        {
          std::thread th;
          std::mutex m;
          std::condition_variable cv;
          std::function<void()> job;
          bool has_job = false, quit = false, done = true;
          std::exception_ptr err;
        };
        std::vector<std::unique_ptr<worker>> w;

        static void loop(worker *k)
        {
          for (;;)
          {
            std::function<void()> job;
            {
              std::unique_lock<std::mutex> lk(k->m);
              k->cv.wait(lk, [&] { return k->has_job || k->quit; });
              if (k->quit) return;
              job.swap(k->job);
              k->has_job = false;
            }
            std::exception_ptr err;
            try { job(); } catch (...) { err = std::current_exception(); }
            {
              std::lock_guard<std::mutex> lk(k->m);
              k->err = err;
              k->done = true;
            }
            k->cv.notify_all();
          }
        }

        public:
        explicit slab_workers(int n)
        {
          for (int i = 0; i < n; ++i)
          {
            w.emplace_back(new worker);
            w.back()->th = std::thread(loop, w.back().get());
          }
        }
        ~slab_workers()
        {
          for (auto &k : w)
          {
            { std::lock_guard<std::mutex> lk(k->m); k->quit = true; }
            k->cv.notify_all();
            k->th.join();
          }
        }
        // runs f(d) for every slab d on that slab's thread; returns when all are done (= a barrier between phases);
        // the first exception is re-thrown in the caller
        template <class F>
        void run(F f)
        {
          for (size_t d = 0; d < w.size(); ++d)
          {
            worker *k = w[d].get();
            { std::lock_guard<std::mutex> lk(k->m); k->job = [f, d] { f(int(d)); }; k->has_job = true; k->done = false; k->err = nullptr; }
            k->cv.notify_all();
          }
          std::exception_ptr first;
          for (auto &k : w)
          {
            std::unique_lock<std::mutex> lk(k->m);
            k->cv.wait(lk, [&] { return k->done; });
            if (k->err && !first) first = k->err;
          }
          if (first) std::rethrow_exception(first);
        }
      };

      // ======================================================================================================
      // one x-slab on one device: the reference's particles_t<real_t, CUDA>
      // ======================================================================================================
      template <class real_t>
      struct slab
      {
        typedef unsigned long long n_t;
        typedef lcx_api<real_t> L;         // the engine of this precision: liblcx_b200.so (double) or liblcx_b200_f32.so (float)
        static void chk(int rc) { if (rc != 0) throw std::runtime_error(std::string("libcloudph++ (B200 engine): ") + L::last_error()); }

        opts_init_t<real_t> oi;            // this slab's options
        typename L::engine *e = nullptr;
        int n_dims;
        size_t n_cell;
        int n_x_bfr = 0, n_x_tot = 0;
        size_t n_cell_bfr = 0;
        int halo_size, halo_x, halo_y, halo_z;
        std::pair<int, int> bcond;
        real_t lft_x1 = -1, rgt_x0 = -1;
        bool spawned = false;              // part of a multi-slab run: migration and post_copy are driven from outside

        bool init_called = false, should_now_run_async = false, should_now_run_cond = false, var_rho = false;
        real_t dt;
        int sstp_cond, sstp_coal, sstp_cond_act;
        as_t adve_scheme;
        bool allow_sstp_cond, pure_const_multi;

        int rng_mode;
        std::mt19937 engine;               // one generator per object, as src/detail/urand.hpp:24,57
        uint64_t philox_call = 0;
        int slab_rank = 0;                 // position of this slab in the decomposition: second Philox key word
        size_t philox_cell_base = 0;       // process-distributed runs: cells of the ranks before this one (n_cell_bfr stays 0 there)
        std::vector<uint32_t> un_host;
        std::vector<real_t> u01_host;
        // caller-supplied streams for the next coalescence sub-steps (lgrngn_b200_inject_rng), consumed before anything is drawn
        std::deque<std::pair<std::vector<uint32_t>, std::vector<real_t>>> injected;
        bool dense_sid;

        // l2e[q] = element of the caller's array that feeds cell q; runs = the same map as maximal contiguous pieces.
        // Few runs (contiguous arrays: 1 for scalars, 3 for a Courant field with its periodic halo) are copied straight
        // between the caller's memory and the device; anything else goes through a page-locked staging buffer.
        struct run_t { long dst, src, len; };
        struct map_t
        {
          std::vector<long> l2e;
          std::vector<run_t> runs;
          real_t *pinned = nullptr;
          bool direct() const { return runs.size() <= 16; }
          void find_runs()
          {
            runs.clear();
            for (size_t q = 0; q < l2e.size(); ++q)
            {
              if (!runs.empty() && l2e[q] == runs.back().src + runs.back().len) ++runs.back().len;
              else runs.push_back(run_t{long(q), l2e[q], 1});
            }
          }
          ~map_t() { if (pinned) L::host_free(pinned); }
        };
        map_t m_th, m_rv, m_rhod, m_p, m_cx, m_cy, m_cz;
        std::vector<real_t> outbuf_host;
        std::vector<std::pair<const map_t *, real_t *>> pending_out;
        bool transfers_open = false;       // uploads of a deferred sync_in not yet waited for
        std::map<common::output_t, real_t> puddle0;

        slab(const opts_init_t<real_t> &o, std::pair<int, int> bc, int n_x_tot_) : oi(o), bcond(bc)
        {
          oi.dev_count = 0;
          n_dims = oi.nx / m1(oi.nx) + oi.ny / m1(oi.ny) + oi.nz / m1(oi.nz);
          n_cell = size_t(m1(oi.nx)) * m1(oi.ny) * m1(oi.nz);
          n_x_tot = n_x_tot_;
          halo_size = oi.adve_scheme == as_t::pred_corr ? 2 : 0;
          halo_x = n_dims == 1 ? halo_size : n_dims == 2 ? halo_size * oi.nz : halo_size * oi.nz * oi.ny;
          halo_y = halo_size * (oi.ny + 1) * oi.nz;
          halo_z = n_dims == 2 ? halo_size * (oi.nz + 1) : halo_size * (oi.nz + 1) * oi.ny;
          adve_scheme = oi.adve_scheme;
          sstp_cond = oi.sstp_cond; sstp_coal = oi.sstp_coal; sstp_cond_act = oi.sstp_cond_act; dt = oi.dt;
          allow_sstp_cond = oi.sstp_cond > 1 || oi.sstp_cond_act > 1;
          pure_const_multi = (oi.sd_conc == 0) && (oi.sd_const_multi > 0 || oi.dry_sizes.size() > 0);
          rng_mode = g_rng_mode;
          dense_sid = g_dense_sid < 0 ? rng_mode == LGRNGN_B200_RNG_MT19937 : g_dense_sid != 0;
          engine.seed(oi.rng_seed);
          unsupported_options();
        }

        ~slab() { if (e) L::destroy(e); }

        void unsupported_options() const
        {
          // everything SURVEY.md section 8 marks out of scope is refused loudly instead of being ignored
          if (oi.chem_switch) throw std::runtime_error("libcloudph++: aqueous chemistry (chem_switch) is not part of the B200 back-end");
          if (oi.ice_switch) throw std::runtime_error("libcloudph++: ice microphysics (ice_switch) is not part of the B200 back-end");
          if (oi.turb_adve_switch || oi.turb_cond_switch || oi.turb_coal_switch)
            throw std::runtime_error("libcloudph++: SGS turbulence for super-droplets (turb_*_switch) is not part of the B200 back-end");
          if (oi.src_type != src_t::off) throw std::runtime_error("libcloudph++: aerosol sources (src_type) are not part of the B200 back-end");
          if (oi.rlx_switch) throw std::runtime_error("libcloudph++: aerosol relaxation (rlx_switch) is not part of the B200 back-end");
          if (oi.diag_incloud_time) throw std::runtime_error("libcloudph++: diag_incloud_time is not part of the B200 back-end");
          if (oi.kernel == kernel_t::onishi_hall || oi.kernel == kernel_t::onishi_hall_davis_no_waals)
            throw std::runtime_error("libcloudph++: To use the turbulent Onishis kernel, set turb_coal_switch=True");
        }

        bool is_distmem() const { return bcond.first == distmem || bcond.second == distmem; }

        // ---- sanity checks of init(): init_sanity_check.ipp:14-176 (the applicable subset, same texts) --------
        void init_sanity_check(const arrinfo_t<real_t> &th, const arrinfo_t<real_t> &rv, const arrinfo_t<real_t> &rhod,
                               const arrinfo_t<real_t> &p, const arrinfo_t<real_t> &cx, const arrinfo_t<real_t> &cy,
                               const arrinfo_t<real_t> &cz, size_t n_ambient_chem)
        {
          if (init_called) throw std::runtime_error("libcloudph++: init() may be called just once");
          init_called = true;
          if (th.is_null() || rv.is_null() || rhod.is_null())
            throw std::runtime_error("libcloudph++: passing th, rv and rhod is mandatory");
          check_courants(cx, cy, cz);
          if (n_ambient_chem != 0) throw std::runtime_error("libcloudph++: chemistry was switched off and ambient_chem is not empty");
          if (oi.dry_distros.size() == 0 && oi.dry_sizes.size() == 0)
            throw std::runtime_error("libcloudph++: Both dry_distros and dry_sizes are undefined");
          if (n_dims > 0)
          {
            if (!(oi.x0 >= 0 && oi.x0 < m1(oi.nx) * oi.dx)) throw std::runtime_error("libcloudph++: !(x0 >= 0 & x0 < min(1,nx)*dz)");
            if (!(oi.y0 >= 0 && oi.y0 < m1(oi.ny) * oi.dy)) throw std::runtime_error("libcloudph++: !(y0 >= 0 & y0 < min(1,ny)*dy)");
            if (!(oi.z0 >= 0 && oi.z0 < m1(oi.nz) * oi.dz)) throw std::runtime_error("libcloudph++: !(z0 >= 0 & z0 < min(1,nz)*dz)");
            if (!(oi.y1 > oi.y0 && oi.y1 <= m1(oi.ny) * oi.dy)) throw std::runtime_error("libcloudph++: !(y1 > y0 & y1 <= min(1,ny)*dy)");
            if (!(oi.z1 > oi.z0 && oi.z1 <= m1(oi.nz) * oi.dz)) throw std::runtime_error("libcloudph++: !(z1 > z0 & z1 <= min(1,nz)*dz)");
          }
          if (!oi.aerosol_conc_factor.empty() && n_dims < 2)
            throw std::runtime_error("libcloudph++: aerosol_conc_factor can only be used in 2D and 3D");
          if (!oi.aerosol_conc_factor.empty() && size_t(oi.nz) != oi.aerosol_conc_factor.size())
            throw std::runtime_error("libcloudph++: aerosol_conc_factor size needs to be either 0 or nz");
          if (!oi.aerosol_conc_factor.empty() && !oi.aerosol_independent_of_rhod)
            throw std::runtime_error("libcloudph++: aerosol_conc_factor can only be used if aerosol_independent_of_rhod==true");
          if (oi.dt == 0) throw std::runtime_error("libcloudph++: please specify opts_init.dt");
          if (oi.sd_conc * oi.sd_const_multi != 0)
            throw std::runtime_error("libcloudph++: specify either opts_init.sd_conc or opts_init.sd_const_multi, not both");
          if (oi.sd_conc == 0 && oi.sd_const_multi == 0 && oi.dry_sizes.size() == 0)
            throw std::runtime_error("libcloudph++: please specify opts_init.sd_conc, opts_init.sd_const_multi or opts_init.dry_sizes");
          if (oi.coal_switch)
          {
            if (oi.terminal_velocity == vt_t::undefined)
              throw std::runtime_error("libcloudph++: please specify opts_init.terminal_velocity or turn off opts_init.coal_switch");
            if (oi.kernel == kernel_t::undefined) throw std::runtime_error("libcloudph++: please specify opts_init.kernel");
          }
          if (oi.sedi_switch && oi.terminal_velocity == vt_t::undefined)
            throw std::runtime_error("libcloudph++: please specify opts_init.terminal_velocity or turn off opts_init.sedi_switch");
          if (oi.sedi_switch && oi.nz == 0) throw std::runtime_error("libcloudph++: opts_init.sedi_switch can be True only if n_dims > 1");
          if (oi.subs_switch && oi.nz == 0) throw std::runtime_error("libcloudph++: opts_init.subs_switch can be True only if n_dims > 1");
          if (oi.subs_switch && size_t(oi.nz) != oi.w_LS.size())
            throw std::runtime_error("libcloudph++: opts_init.subs_switch == True, but subsidence velocity profile size != nz");
          if (oi.const_p && p.is_null())
            throw std::runtime_error("libcloudph++: In const_p option, pressure profile must be passed (p in init())");
          if (!oi.const_p && !p.is_null())
            throw std::runtime_error("libcloudph++: pressure profile was passed in init(), but the constant pressure option was not used");
          if (oi.sstp_cond < 1) throw std::runtime_error("libcloudph++: opts_init.sstp_cond needs to be greater than 0");
          if (oi.adaptive_sstp_cond && !oi.exact_sstp_cond)
            throw std::runtime_error("libcloudph++: Adaptive condensation substepping (opts_init.adaptive_sstp_cond) works oly for per-particle substepping (opts_init.exact_sstp_cond)");
          if (!oi.sstp_cond_mix && !oi.exact_sstp_cond)
            throw std::runtime_error("libcloudph++: Mixing of rv and th (opts_init.sstp_cond_mix) can only be disable for per-particle substepping (opts_init.exact_sstp_cond)");
          if (oi.sstp_cond_mix && oi.adaptive_sstp_cond && oi.exact_sstp_cond)
            throw std::runtime_error("libcloudph++: Adaptive cond substepping (opts_init.adaptive_sstp_cond) with per-particle substepping (opts_init.exact_sstp_cond) requires mixing of th and rv between subteps (opts_init.sstp_cond_mix) to be disabled");
          if (oi.sstp_cond_act > 1 && (oi.sstp_cond_mix || !oi.exact_sstp_cond || !oi.adaptive_sstp_cond))
            throw std::runtime_error("libcloudph++: number of substeps for activation (opts_init.sstp_cond_act) can be greater than 1 only if mixing of rv and th (opts_init.sstp_cond_mix) is disabled and if per-particle condensation substepping is used (opts_init.exact_sstp_cond) and if adaptive substepping is used (opts_init.adaptive_sstp_cond)");
        }

        void check_courants(const arrinfo_t<real_t> &cx, const arrinfo_t<real_t> &cy, const arrinfo_t<real_t> &cz) const
        {
          if (!cx.is_null() || !cy.is_null() || !cz.is_null())
          {
            if (n_dims == 0) throw std::runtime_error("libcloudph++: Courant numbers passed in 0D setup");
            if (n_dims == 1 && (cx.is_null() || !cy.is_null() || !cz.is_null()))
              throw std::runtime_error("libcloudph++: Only X Courant number allowed in 1D setup");
            if (n_dims == 2 && (cx.is_null() || !cy.is_null() || cz.is_null()))
              throw std::runtime_error("libcloudph++: Only X and Z Courant numbers allowed in 2D setup");
            if (n_dims == 3 && (cx.is_null() || cy.is_null() || cz.is_null()))
              throw std::runtime_error("libcloudph++: All XYZ Courant number components required in 3D setup");
          }
        }

        // ---- Eulerian <-> Lagrangian index map: init_e2l.ipp:34-114 ---------------------------------------------
        void init_e2l(const arrinfo_t<real_t> &arr, map_t &m, int field, int ext_x = 0, int ext_y = 0, int ext_z = 0, long offset = 0)
        {
          int64_t count = 0;
          chk(L::field_size(e, field, &count));
          m.l2e.resize(size_t(count));
          const long shift = long(n_cell_bfr) + offset;
          long max_stride = 0, max_stride_n_cell = 0;
          switch (n_dims)
          {
            case 0: m.l2e[0] = 0; break;
            case 1:
              for (long q = 0; q < long(count); ++q) m.l2e[size_t(q)] = int(shift + q);
              max_stride = 1; max_stride_n_cell = n_x_tot + ext_x;
              break;
            case 2:
              for (long q = 0; q < long(count); ++q)
              {
                const int v = int(shift + q);
                m.l2e[size_t(q)] = arr.strides[0] * (v / (oi.nz + ext_z)) + arr.strides[1] * (v % (oi.nz + ext_z));
              }
              max_stride = arr.strides[0]; max_stride_n_cell = n_x_tot + ext_x;
              break;
            case 3:
              for (long q = 0; q < long(count); ++q)
              {
                const int v = int(shift + q);
                m.l2e[size_t(q)] = arr.strides[0] * (v / ((oi.nz + ext_z) * (oi.ny + ext_y)))
                                 + arr.strides[1] * ((v / (oi.nz + ext_z)) % (oi.ny + ext_y))
                                 + arr.strides[2] * (v % (oi.nz + ext_z));
              }
              max_stride = std::max(arr.strides[0], arr.strides[1]);
              max_stride_n_cell = max_stride == arr.strides[0] ? n_x_tot + ext_x : oi.ny + ext_y;
              break;
          }
          const long n_tot = max_stride_n_cell * max_stride;
          if (n_dims > 0)
            for (long &l : m.l2e) { if (l >= n_tot) l -= n_tot; else if (l < 0) l += n_tot; }
          m.find_runs();
          if (!m.direct() && !m.pinned)
          {
            void *ptr = nullptr;
            chk(L::host_alloc(m.l2e.size() * sizeof(real_t), &ptr));
            m.pinned = static_cast<real_t *>(ptr);
          }
        }

        // device-pointer fast path of arrinfo_t: a model that keeps its fields on the GPU hands device pointers; copies then stay on the device
        static bool on_device(const void *p) { int d = 0; chk(L::pointer_on_device(p, &d)); return d != 0; }

        // host copy of a caller's array that may live in device memory (initialisation and the Courant-range check read fields on the host)
        struct host_view
        {
          std::vector<real_t> buf;
          const real_t *data;
          host_view(const arrinfo_t<real_t> &a, const map_t &m) : data(a.data)
          {
            if (a.is_null() || !on_device(a.data)) return;
            long hi = 0;
            for (const run_t &r : m.runs) hi = std::max(hi, r.src + r.len);
            buf.resize(size_t(hi));
            chk(L::copy_to_host(buf.data(), a.data, buf.size() * sizeof(real_t)));
            data = buf.data();
          }
        };

        // both directions only queue the copies; the caller ends the batch with finish_transfers()
        void sync_in_field(const arrinfo_t<real_t> &from, const map_t &m, int field)   // impl_sync.ipp:15-40
        {
          if (from.is_null()) return;
          const real_t *src = from.data;
          if (!m.direct() && on_device(src))
            throw std::runtime_error("libcloudph++ (B200 engine): Eulerian arrays in device memory must be (nearly) contiguous: z fastest, no padding between rows");
          if (m.direct())
          {
            for (const run_t &r : m.runs) chk(L::cells_set_part(e, field, r.dst, src + r.src, r.len));
            return;
          }
          const long n = long(m.l2e.size());
#pragma omp parallel for schedule(static)
          for (long q = 0; q < n; ++q) m.pinned[q] = src[m.l2e[size_t(q)]];
          chk(L::cells_set_part(e, field, 0, m.pinned, n));
        }
        void sync_out_field(int field, const map_t &m, arrinfo_t<real_t> &to)           // impl_sync.ipp:42-68
        {
          if (to.is_null()) return;
          if (!m.direct() && on_device(to.data))
            throw std::runtime_error("libcloudph++ (B200 engine): Eulerian arrays in device memory must be (nearly) contiguous: z fastest, no padding between rows");
          if (m.direct())
          {
            for (const run_t &r : m.runs) chk(L::cells_get_part(e, field, r.dst, to.data + r.src, r.len));
            return;
          }
          chk(L::cells_get_part(e, field, 0, m.pinned, int64_t(m.l2e.size())));
          pending_out.push_back(std::make_pair(&m, to.data));
        }
        void finish_transfers()
        {
          chk(L::sync(e));
          for (const auto &po : pending_out)
          {
            const map_t &m = *po.first;
            real_t *dst = po.second;
            const long n = long(m.l2e.size());
#pragma omp parallel for schedule(static)
            for (long q = 0; q < n; ++q) dst[m.l2e[size_t(q)]] = m.pinned[q];
          }
          pending_out.clear();
        }

        // ---- engine creation ------------------------------------------------------------------------------------
        void create_engine()
        {
          lcx_config c;
          std::memset(&c, 0, sizeof(c));
          c.device = oi.dev_id;
          c.real_bytes = sizeof(real_t);
          c.nx = oi.nx; c.ny = oi.ny; c.nz = oi.nz;
          c.dx = oi.dx; c.dy = oi.dy; c.dz = oi.dz;
          c.x0 = oi.x0; c.y0 = oi.y0; c.z0 = oi.z0; c.x1 = oi.x1; c.y1 = oi.y1; c.z1 = oi.z1;
          c.n_sd_max = oi.n_sd_max;
          c.kernel = int(oi.kernel); c.terminal_velocity = int(oi.terminal_velocity);
          c.adve_scheme = int(oi.adve_scheme); c.RH_formula = int(oi.RH_formula);
          c.th_dry = oi.th_dry; c.const_p = oi.const_p;
          c.n_kernel_user_params = int(std::min<size_t>(oi.kernel_parameters.size(), 4));
          for (int q = 0; q < c.n_kernel_user_params; ++q) c.kernel_user_params[q] = oi.kernel_parameters[size_t(q)];
          c.open_side_walls = oi.open_side_walls; c.periodic_topbot_walls = oi.periodic_topbot_walls;
          c.bcond_lft = bcond.first; c.bcond_rgt = bcond.second;
          c.lft_x1 = lft_x1; c.rgt_x0 = rgt_x0;
          c.multi_kappa = (oi.dry_distros.size() + oi.dry_sizes.size() > 1);
          c.pure_const_multi = pure_const_multi;
          c.allow_sstp_cond = allow_sstp_cond;
          c.exact_sstp_cond = oi.exact_sstp_cond;
          c.sstp_cond_act = oi.sstp_cond_act;
          c.rc2_T = oi.rc2_T;
          std::vector<real_t> eff;
          if (oi.coal_switch) eff = init_kernel(c);
          chk(L::create(&c, &e));
          chk(L::set_dense_storage_index(e, dense_sid));
          if (!eff.empty()) chk(L::set_efficiencies(e, eff.data(), int64_t(eff.size())));
        }

        // kernel parameter checks and efficiency tables: init_kernel.ipp:6-235
        std::vector<real_t> init_kernel(lcx_config &c) const
        {
          const size_t n_user = oi.kernel_parameters.size();
          std::vector<real_t> eff;
          auto table = [&](const char *name, const char *what, double r_max) {
            if (n_user != 0) throw std::runtime_error(std::string("libcloudph++: ") + what + " kernel doesn't accept parameters.");
            eff = load_efficiencies<real_t>(name);
            c.kernel_r_max = r_max;
          };
          switch (oi.kernel)
          {
            case kernel_t::golovin:
              if (n_user != 1) throw std::runtime_error("libcloudph++: Golovin kernel accepts exactly one parameter.");
              break;
            case kernel_t::geometric:
              if (n_user > 1) throw std::runtime_error("libcloudph++: Geometric kernel accepts up to one parameter.");
              break;
            case kernel_t::Long:
              if (n_user > 0) throw std::runtime_error("libcloudph++: Long kernel doesn't take parameters.");
              break;
            case kernel_t::hall:                      table("hall", "Hall", 300.); break;
            case kernel_t::hall_davis_no_waals:       table("hall_davis_no_waals", "Hall + Davis", 1100.); break;
            case kernel_t::vohl_davis_no_waals:       table("vohl_davis_no_waals", "Vohl + Davis", 0.); break;
            case kernel_t::hall_pinsky_stratocumulus: table("hall_pinsky_stratocumulus", "Hall + Pinsky (stratocumulus)", 0.); break;
            case kernel_t::hall_pinsky_1000mb_grav:   table("hall_pinsky_1000mb_grav", "Hall + Pinsky (gravitational 1000mb)", 0.); break;
            case kernel_t::hall_pinsky_cumulonimbus:  table("hall_pinsky_cumulonimbus", "Hall + Pinsky (cumulonimbus)", 0.); break;
            default: break;
          }
          if (!eff.empty())
          {
            // the first value of each data file is r_max [um] of that table (written by tools/extract_efficiencies.py)
            c.kernel_r_max = double(eff.front());
            eff.erase(eff.begin());
          }
          return eff;
        }

        // ---- init(): particles_init.ipp:16-131 -------------------------------------------------------------------
        void init(const arrinfo_t<real_t> &th, const arrinfo_t<real_t> &rv, const arrinfo_t<real_t> &rhod, const arrinfo_t<real_t> &p,
                  const arrinfo_t<real_t> &cx, const arrinfo_t<real_t> &cy, const arrinfo_t<real_t> &cz, size_t n_ambient_chem)
        {
          stopwatch sw;
          init_sanity_check(th, rv, rhod, p, cx, cy, cz, n_ambient_chem);
          if (oi.rng_seed_init_switch) engine.seed(oi.rng_seed_init);
          create_engine();
          sw.lap("create_engine");

          init_e2l(th, m_th, LCX_F_TH);
          init_e2l(rv, m_rv, LCX_F_RV);
          init_e2l(rhod, m_rhod, LCX_F_RHOD);
          if (oi.const_p) init_e2l(p, m_p, LCX_F_P);
          init_courant_maps(cx, cy, cz);
          sw.lap("index maps");

          sync_in_field(th, m_th, LCX_F_TH);
          sync_in_field(rv, m_rv, LCX_F_RV);
          sync_in_field(rhod, m_rhod, LCX_F_RHOD);
          if (oi.const_p) sync_in_field(p, m_p, LCX_F_P);
          sync_in_field(cx, m_cx, LCX_F_COURANT_X);
          sync_in_field(cy, m_cy, LCX_F_COURANT_Y);
          sync_in_field(cz, m_cz, LCX_F_COURANT_Z);
          finish_transfers();
          if (oi.subs_switch) chk(L::cells_set(e, LCX_F_W_LS, oi.w_LS.data(), int64_t(oi.w_LS.size()), 0));

          chk(L::hskpng_Tpr(e));
          sw.lap("field uploads");

          if (!oi.no_ccn_at_init)
          {
            const cell_state cs = host_cells(th, rv, rhod, p);
            sw.lap("host cell state");
            if (oi.dry_distros.size() > 0) init_SD_with_distros(cs);
            if (oi.dry_sizes.size() > 0) init_SD_with_sizes(cs);
            sw.lap("super-droplets (total)");
          }

          if (oi.terminal_velocity == vt_t::beard77fast)
          {
            // cached sea-level fall speeds, evaluated on the host like the reference's CPU back-ends: init_vterm.ipp:36-59
            const lcx::vt0_bins<real_t> bins;
            std::vector<real_t> vt0(lcx::VT0_N_BIN);
            for (int it = 0; it < lcx::VT0_N_BIN; ++it) vt0[size_t(it)] = lcx::vt_beard77_v0(bins.mid(it));
            chk(L::set_vt0_table(e, vt0.data(), int(vt0.size())));
          }
          chk(L::hskpng_vterm(e, 1));
          chk(L::hskpng_rc2(e));                        // critical radii for activation sub-stepping (particles_init.ipp:116-117)
          chk(L::sstp_save(e));
          chk(L::post_copy(e, 0, /*keep_all=*/1));      // hskpng_count(): group by the cells assigned at creation
          chk(L::sync(e));
          sw.lap("vterm, rc2, grouping");
          engine.seed(oi.rng_seed);
          philox_call = 0;
        }

        void init_courant_maps(const arrinfo_t<real_t> &cx, const arrinfo_t<real_t> &cy, const arrinfo_t<real_t> &cz)
        {
          if (!cx.is_null()) init_e2l(cx, m_cx, LCX_F_COURANT_X, 1, 0, 0, -halo_x);
          if (!cy.is_null()) init_e2l(cy, m_cy, LCX_F_COURANT_Y, 0, 1, 0, long(n_x_bfr) * oi.nz - halo_y);
          if (!cz.is_null()) init_e2l(cz, m_cz, LCX_F_COURANT_Z, 0, 0, 1, long(n_x_bfr) * std::max(1, oi.ny) - halo_z);
        }

        // host copies of the per-cell state the initialisation needs, computed with the host libm
        struct cell_state { std::vector<real_t> rhod, T, RH, dv; };

        cell_state host_cells(const arrinfo_t<real_t> &th, const arrinfo_t<real_t> &rv, const arrinfo_t<real_t> &rhod, const arrinfo_t<real_t> &p) const
        {
          cell_state cs;
          cs.rhod.resize(n_cell); cs.T.resize(n_cell); cs.RH.resize(n_cell); cs.dv.resize(n_cell);
          const host_view h_th(th, m_th), h_rv(rv, m_rv), h_rhod(rhod, m_rhod), h_p(p, m_p);
#pragma omp parallel for schedule(static)
          for (long cl = 0; cl < long(n_cell); ++cl)
          {
            const size_t c = size_t(cl);
            const real_t th_c = h_th.data[m_th.l2e[c]], rv_c = h_rv.data[m_rv.l2e[c]], rhod_c = h_rhod.data[m_rhod.l2e[c]];
            real_t T_c, p_c;
            if (oi.th_dry) T_c = lcx::T_of_th_dry(th_c, rhod_c);
            else           T_c = th_c * lcx::exner(h_p.data[m_p.l2e[c]]);
            p_c = oi.const_p ? h_p.data[m_p.l2e[c]] : lcx::p_of_rhod_rv_T(rhod_c, rv_c, T_c);
            cs.rhod[c] = rhod_c; cs.T[c] = T_c;
            cs.RH[c] = lcx::RH_of(int(oi.RH_formula), p_c, rv_c, T_c);
            if (n_dims == 0) cs.dv[c] = real_t(1) / rhod_c;
            else
            {
              const int ic = int(c), nz1 = std::max(1, oi.nz), ny1 = std::max(1, oi.ny);
              const int i = (ic / nz1) / ny1, j = (ic / nz1) % ny1, k = ic % nz1;
              cs.dv[c] = std::max(real_t(0),
                (std::min((i + 1) * oi.dx, oi.x1) - std::max(i * oi.dx, oi.x0)) *
                (std::min((j + 1) * oi.dy, oi.y1) - std::max(j * oi.dy, oi.y0)) *
                (std::min((k + 1) * oi.dz, oi.z1) - std::max(k * oi.dz, oi.z0)));
            }
          }
          return cs;
        }

        // range of ln(rd) that the super-droplets of one spectrum must cover: init_dist_analysis.ipp:17-75
        void dist_analysis_sd_conc(const common::unary_function<real_t> &fun, const cell_state &cs,
                                   real_t &log_rd_min, real_t &log_rd_max, real_t &multiplier) const
        {
          const n_t sd_conc = oi.sd_conc;
          const real_t dt_ = 1;
          const real_t vol = n_dims == 0 ? cs.dv[0] : (oi.dx * oi.dy * oi.dz);
          if (oi.rd_min >= 0 && oi.rd_max >= 0)
          {
            const real_t rd_min = oi.rd_min, rd_max = oi.rd_max;
            multiplier = std::log(rd_max / rd_min) / sd_conc * dt_ * vol;
            log_rd_min = std::log(rd_min);
            log_rd_max = std::log(rd_max);
            return;
          }
          if (!(oi.rd_min < 0 && oi.rd_max < 0)) throw std::runtime_error("libcloudph++: opts_init.rd_min * opts_init.rd_max < 0");
          const real_t rd_min_init = 1e-14, rd_max_init = 1e-3;      // src/detail/config.hpp:23-24
          real_t rd_min = rd_min_init, rd_max = rd_max_init;
          bool found = false;
          while (!found)
          {
            multiplier = std::log(rd_max / rd_min) / sd_conc * dt_ * vol;
            log_rd_min = std::log(rd_min);
            log_rd_max = std::log(rd_max);
            const n_t n_min = n_t(fun(log_rd_min) * multiplier), n_max = n_t(fun(log_rd_max) * multiplier);
            if (rd_min == rd_min_init && n_min != 0)
            { std::ostringstream s; s << "Initial dry radii distribution is non-zero (" << n_min << ") for rd_min_init (" << rd_min_init << ")"; throw std::runtime_error(s.str()); }
            if (rd_max == rd_max_init && n_max != 0)
            { std::ostringstream s; s << "Initial dry radii distribution is non-zero (" << n_max << ") for rd_max_init (" << rd_max_init << ")"; throw std::runtime_error(s.str()); }
            if (n_min == 0) rd_min *= 1.01;
            else if (n_max == 0) rd_max /= 1.01;
            else found = true;
          }
        }

        real_t draw_u01() { return std::uniform_real_distribution<real_t>(0, 1)(engine); }

        // Super-droplets of the sd_conc flavour are created on the device when the random stream is the counter-based one
        // (the initial state is then another sample of the same distributions than the reference's mt19937 gives - like every
        // later draw in that mode); LCX_DEVICE_INIT=0 keeps the host path, which the replayed mt19937 stream always takes.
        uint64_t init_call = 0;
        bool device_init() const
        {
          if (rng_mode == LGRNGN_B200_RNG_MT19937) return false;
          const char *v = std::getenv("LCX_DEVICE_INIT");
          return !(v && v[0] == '0');
        }

        // LCX_INIT_TIMING=1: phases of init() on stderr (where the seconds of a large initialisation go)
        struct stopwatch
        {
          bool on;
          std::chrono::steady_clock::time_point t;
          stopwatch() : on(std::getenv("LCX_INIT_TIMING") && std::getenv("LCX_INIT_TIMING")[0] == '1'), t(std::chrono::steady_clock::now()) {}
          void lap(const char *what)
          {
            if (!on) return;
            const auto now = std::chrono::steady_clock::now();
            std::fprintf(stderr, "[lcx init] %-28s %8.3f s\n", what, std::chrono::duration<double>(now - t).count());
            t = now;
          }
        };

        // Brent's minimiser (Brent 1973, ch. 5) with the stopping rule and step choices of the Boost.Math routine the
        // reference calls in init_dist_analysis.ipp:95 (brent_find_minima with bits capped at half the mantissa)
        template <class F>
        static std::pair<real_t, real_t> brent_minimum(const F &f, real_t lo, real_t hi, int max_iter)
        {
          const real_t tol = std::ldexp(real_t(1), 1 - std::numeric_limits<real_t>::digits / 2);
          const real_t golden = real_t(0.3819660f);
          real_t x = hi, w = hi, v = hi, fx = f(x), fw = fx, fv = fx, d = 0, d_prev = 0;
          for (int it = 0; it < max_iter; ++it)
          {
            const real_t mid = (lo + hi) / 2, t1 = tol * std::fabs(x) + tol / 4, t2 = 2 * t1;
            if (std::fabs(x - mid) <= t2 - (hi - lo) / 2) break;
            bool parabolic = false;
            if (std::fabs(d_prev) > t1)
            {
              real_t r = (x - w) * (fx - fv), q = (x - v) * (fx - fw), pnum = (x - v) * q - (x - w) * r;
              q = 2 * (q - r);
              if (q > 0) pnum = -pnum;
              q = std::fabs(q);
              const real_t d_before = d_prev;
              d_prev = d;
              parabolic = !(std::fabs(pnum) >= std::fabs(q * d_before / 2) || pnum <= q * (lo - x) || pnum >= q * (hi - x));
              if (parabolic)
              {
                d = pnum / q;
                const real_t u = x + d;
                if ((u - lo) < t2 || (hi - u) < t2) d = (mid - x) < 0 ? -std::fabs(t1) : std::fabs(t1);
              }
            }
            if (!parabolic)
            {
              d_prev = (x >= mid) ? lo - x : hi - x;
              d = golden * d_prev;
            }
            const real_t u = std::fabs(d) >= t1 ? x + d : (d > 0 ? x + std::fabs(t1) : x - std::fabs(t1));
            const real_t fu = f(u);
            if (fu <= fx)
            {
              (u >= x ? lo : hi) = x;
              v = w; fv = fw; w = x; fw = fx; x = u; fx = fu;
            }
            else
            {
              (u < x ? lo : hi) = u;
              if (fu <= fw || w == x) { v = w; fv = fw; w = u; fw = fu; }
              else if (fu <= fv || v == x || v == w) { v = u; fv = fu; }
            }
          }
          return std::make_pair(x, fx);
        }

        // ln(rd) range of a spectrum for constant-multiplicity sampling: where it drops to 1e-20 of its maximum
        // (init_dist_analysis.ipp:78-123, config.hpp:17-24)
        void dist_analysis_const_multi(const common::unary_function<real_t> &fun, real_t &log_rd_min, real_t &log_rd_max) const
        {
          if (oi.rd_min >= 0 && oi.rd_max >= 0)
          {
            log_rd_min = std::log(oi.rd_min);
            log_rd_max = std::log(oi.rd_max);
            return;
          }
          if (!(oi.rd_min < 0 && oi.rd_max < 0)) throw std::runtime_error("libcloudph++: opts_init.rd_min * opts_init.rd_max < 0");
          const real_t lo = std::log(real_t(1e-14)), hi = std::log(real_t(1e-3)), threshold = 1e20;
          const auto top = brent_minimum([&](real_t x) { return real_t(-1) * fun(x); }, lo, hi, 100);
          const real_t bound = -top.second / threshold;
          const real_t minus_bound = -bound;
          auto g = [&](real_t x) { return minus_bound + fun(x); };
          const lcx::width_tol<real_t> tol(sizeof(real_t) * 8 / 4);
          uintmax_t it = 100;
          log_rd_min = lcx::toms748(g, lo, top.first, g(lo), g(top.first), tol, it);
          it = 100;
          log_rd_max = lcx::toms748(g, top.first, hi, g(top.first), g(hi), tol, it);
        }

        // concentration at STP -> number of particles in each cell: init_count_num.ipp:35-64
        std::vector<real_t> conc_to_number(const cell_state &cs, real_t conc) const
        {
          std::vector<real_t> arr(n_cell, conc);
          const real_t rho_stp = lcx::cst<real_t>::rho_stp();
          for (size_t c = 0; c < n_cell; ++c)
          {
            arr[c] = arr[c] * cs.dv[c];
            if (!oi.aerosol_independent_of_rhod) arr[c] = cs.rhod[c] / rho_stp * arr[c];
            if (!oi.aerosol_conc_factor.empty()) arr[c] = arr[c] * oi.aerosol_conc_factor[c % size_t(oi.nz)];
          }
          return arr;
        }

        // cell of every new SD: count_num[0] times cell 0, count_num[1] times cell 1, ... (init_ijk.ipp:36-52)
        std::vector<uint32_t> cells_of_new(const std::vector<size_t> &count_num) const
        {
          size_t tot = 0;
          for (const size_t c : count_num) tot += c;
          std::vector<uint32_t> ijk;
          ijk.reserve(tot);
          for (size_t c = 0; c < n_cell; ++c) ijk.insert(ijk.end(), count_num[c], uint32_t(c));
          return ijk;
        }

        // common tail of every initialisation flavour: kappa, equilibrium wet radius (init_wet.ipp:18-74), positions
        // uniform within the part of the cell inside the Lagrangian domain (init_xyz.ipp:16-73), upload
        void finalize_new(const cell_state &cs, const std::vector<uint32_t> &ijk, const std::vector<n_t> &n, const std::vector<real_t> &rd3, real_t kappa)
        {
          const size_t n_new = ijk.size();
          if (n_new == 0) return;
          stopwatch sw;
          std::vector<real_t> rw2(n_new), kpa(n_new, kappa), xs, ys, zs, u01(n_new);
#pragma omp parallel for schedule(static)
          for (long sl = 0; sl < long(n_new); ++sl)
          {
            const size_t s = size_t(sl);
            const real_t RH = std::min(cs.RH[ijk[s]], oi.RH_max);
            rw2[s] = std::pow(lcx::rw3_eq(rd3[s], kpa[s], RH, cs.T[ijk[s]]), real_t(2. / 3));
          }
          sw.lap("  equilibrium wet radii");
          const int nn[3] = {oi.nx, oi.ny, oi.nz};
          const real_t a[3] = {oi.x0, oi.y0, oi.z0}, b[3] = {oi.x1, oi.y1, oi.z1}, d[3] = {oi.dx, oi.dy, oi.dz};
          std::vector<real_t> *v[3] = {&xs, &ys, &zs};
          for (int ix = 0; ix < 3; ++ix)
          {
            if (nn[ix] == 0) continue;
            v[ix]->resize(n_new);
            for (size_t s = 0; s < n_new; ++s) u01[s] = draw_u01();      // the one sequential part: std::mt19937 in the reference's order
            sw.lap("  mt19937 draws (positions)");
            const size_t nz1 = size_t(m1(oi.nz)), ny1 = size_t(m1(oi.ny));
            real_t *out = v[ix]->data();
#pragma omp parallel for schedule(static)
            for (long sl = 0; sl < long(n_new); ++sl)
            {
              const size_t s = size_t(sl);
              const size_t c = ijk[s];
              size_t ii;
              if (n_dims == 1) ii = c;
              else if (n_dims == 2) ii = ix == 0 ? c / nz1 : c % nz1;
              else ii = ix == 0 ? c / (nz1 * ny1) : ix == 1 ? (c / nz1) % ny1 : c % nz1;
              out[s] = u01[s] * std::min(b[ix], (ii + 1) * d[ix]) + (1. - u01[s]) * std::max(a[ix], ii * d[ix]);
            }
            sw.lap("  positions");
          }
          chk(L::sd_append(e, int64_t(n_new), reinterpret_cast<const uint64_t *>(n.data()), rd3.data(), rw2.data(), kpa.data(),
                            xs.empty() ? nullptr : xs.data(), ys.empty() ? nullptr : ys.data(), zs.empty() ? nullptr : zs.data(), ijk.data()));
          sw.lap("  upload (sd_append)");
        }

        // sd_conc SDs per cell, stratified in ln(rd): init_SD_with_distros_sd_conc.ipp:16-52, init_dry_sd_conc.ipp:25-66, init_n.ipp:48-137
        void init_sd_conc(const cell_state &cs, const kappa_rd_insol_t<real_t> &kr, const common::unary_function<real_t> &fun,
                          real_t tot_lnrd_rng, real_t &log_rd_max_out)
        {
          real_t log_rd_min = 0, log_rd_max = 0, multiplier = 0;
          dist_analysis_sd_conc(fun, cs, log_rd_min, log_rd_max, multiplier);
          log_rd_max_out = log_rd_max;
          if (log_rd_min >= log_rd_max)
          { std::ostringstream s; s << "Distribution analysis error: rd_min(" << std::exp(log_rd_min) << ") >= rd_max(" << std::exp(log_rd_max) << ")"; throw std::runtime_error(s.str()); }

          const real_t rho_stp = lcx::cst<real_t>::rho_stp();
          const real_t fraction = (log_rd_max - log_rd_min) / tot_lnrd_rng;
          multiplier *= oi.sd_conc / int(fraction * oi.sd_conc + 0.5);
          const size_t per_cell = size_t(fraction * oi.sd_conc);                      // init_count_num.ipp:32-35
          const std::vector<uint32_t> ijk = cells_of_new(std::vector<size_t>(n_cell, per_cell));
          const size_t n_new = ijk.size();

          stopwatch sw;
          // multiplicity of one SD from the caller's spectrum: init_n.ipp:48-137 (evaluated on the host, like the reference)
          auto multiplicity = [&](size_t s, real_t rd3_s) {
            const real_t lnrd2 = std::log(rd3_s) / 3.;
            real_t v = multiplier * fun(lnrd2);
            if (!oi.aerosol_independent_of_rhod) v = v * cs.rhod[ijk[s]] / rho_stp;
            if (!oi.aerosol_conc_factor.empty()) v = v * oi.aerosol_conc_factor[ijk[s] % size_t(oi.nz)];
            if (n_dims > 0) v = v * cs.dv[ijk[s]] / real_t(oi.dx * oi.dy * oi.dz);
            return n_t(v + real_t(0.5));
          };
          if (device_init())
          {
            // counter-based random stream: dry radii, equilibrium wet radii and positions are made on the device (lcx_init.cu);
            // only the spectrum - an arbitrary functor of the caller's - is evaluated here
            std::vector<real_t> rd3(n_new);
            std::vector<n_t> n(n_new);
            int64_t first = 0;
            chk(L::n_part(e, &first));
            chk(L::sd_append_sd_conc(e, int64_t(per_cell), log_rd_min, log_rd_max, kr.kappa, oi.RH_max, uint64_t(uint32_t(oi.rng_seed)),
                                     uint32_t(slab_rank), ~uint64_t(0) - init_call++, rd3.data()));
            sw.lap("  device: rd3, rw2, xyz");
#pragma omp parallel for schedule(static)
            for (long sl = 0; sl < long(n_new); ++sl) n[size_t(sl)] = multiplicity(size_t(sl), rd3[size_t(sl)]);
            sw.lap("  multiplicities (host)");
            chk(L::sd_set_n(e, first, int64_t(n_new), reinterpret_cast<const uint64_t *>(n.data())));
            sw.lap("  upload of n");
            return;
          }
          std::vector<n_t> n(n_new);
          std::vector<real_t> rd3(n_new), u01(n_new);
          for (size_t s = 0; s < n_new; ++s) u01[s] = draw_u01();
          sw.lap("  mt19937 draws (dry radii)");
#pragma omp parallel for schedule(static)
          for (long sl = 0; sl < long(n_new); ++sl)
          {
            const size_t s = size_t(sl);
            const size_t ptr = size_t(ijk[s]) * per_cell;
            const real_t lnrd = log_rd_min + real_t(s - ptr + u01[s]) * (log_rd_max - log_rd_min) / real_t(per_cell);
            rd3[s] = std::exp(3 * lnrd);
            n[s] = multiplicity(s, rd3[s]);
          }
          sw.lap("  dry radii, multiplicities");
          finalize_new(cs, ijk, n, rd3, kr.kappa);
        }

        // SDs of one fixed multiplicity, dry radii drawn from the spectrum's CDF tabulated with bin_precision = 1e-4 in ln(rd):
        // init_SD_with_distros_const_multi.ipp:16-39 / _tail.ipp:16-40, init_count_num.ipp:67-104, init_dry_const_multi.ipp:25-83
        void init_const_multi(const cell_state &cs, const kappa_rd_insol_t<real_t> &kr, const common::unary_function<real_t> &fun,
                              n_t multi, const real_t *log_rd_min_forced)
        {
          const real_t bin = 1e-4;
          real_t log_rd_min = 0, log_rd_max = 0;
          dist_analysis_const_multi(fun, log_rd_min, log_rd_max);
          if (log_rd_min_forced) log_rd_min = *log_rd_min_forced;
          if (log_rd_min >= log_rd_max)
          { std::ostringstream s; s << "Distribution analysis error: rd_min(" << std::exp(log_rd_min) << ") >= rd_max(" << std::exp(log_rd_max) << ")"; throw std::runtime_error(s.str()); }

          // trapezoid integral of the spectrum -> concentration -> SDs per cell
          const int n_bin = int((log_rd_max - log_rd_min) / bin);
          real_t integral = (fun(log_rd_min) + fun(log_rd_max)) / 2.;
          for (int i = 1; i < n_bin; ++i) integral += fun(log_rd_min + i * bin);
          integral = integral * bin;
          const std::vector<real_t> number = conc_to_number(cs, integral);
          std::vector<size_t> count_num(n_cell);
          for (size_t c = 0; c < n_cell; ++c) count_num[c] = size_t(number[c] / multi + real_t(0.5));
          const std::vector<uint32_t> ijk = cells_of_new(count_num);
          const size_t n_new = ijk.size();

          const size_t n_pt = size_t((log_rd_max - log_rd_min) / bin + 1);
          std::vector<real_t> cdf(n_pt);
          for (size_t i = 0; i < n_pt; ++i) cdf[i] = real_t(1) * fun(log_rd_min + bin * i);
          for (size_t i = 1; i < n_pt; ++i) cdf[i] = cdf[i - 1] + cdf[i];
          const real_t total = cdf.back();
          for (size_t i = 0; i < n_pt; ++i) cdf[i] = cdf[i] / total;

          std::vector<real_t> rd3(n_new);
          for (size_t s = 0; s < n_new; ++s)
          {
            const real_t pos = real_t(std::upper_bound(cdf.begin(), cdf.end(), draw_u01()) - cdf.begin());
            rd3[s] = std::exp(3 * (log_rd_min + pos * bin));
          }
          finalize_new(cs, ijk, std::vector<n_t>(n_new, multi), rd3, kr.kappa);
        }

        void init_SD_with_distros(const cell_state &cs)                      // init_SD_with_distros.ipp:15-60
        {
          real_t tot_lnrd_rng = 0.;
          if (oi.sd_conc > 0)
            for (const auto &dd : oi.dry_distros)
            {
              real_t lo = 0, hi = 0, mult = 0;
              dist_analysis_sd_conc(*dd.second, cs, lo, hi, mult);
              tot_lnrd_rng += hi - lo;
            }
          for (const auto &dd : oi.dry_distros)
          {
            if (oi.sd_conc > 0)
            {
              real_t log_rd_max = 0;
              init_sd_conc(cs, dd.first, *dd.second, tot_lnrd_rng, log_rd_max);
              if (oi.sd_conc_large_tail) init_const_multi(cs, dd.first, *dd.second, 1, &log_rd_max);
            }
            if (oi.sd_const_multi > 0) init_const_multi(cs, dd.first, *dd.second, oi.sd_const_multi, nullptr);
          }
        }

        // monodisperse batches: kappa -> { radius -> (STP concentration, SDs per cell) }: init_SD_with_sizes.ipp:16-76
        void init_SD_with_sizes(const cell_state &cs)
        {
          for (const auto &species : oi.dry_sizes)
            for (const auto &size : species.second)
            {
              const real_t radius = size.first, conc = size.second.first;
              const int sd_count = size.second.second;
              const std::vector<uint32_t> ijk = cells_of_new(std::vector<size_t>(n_cell, size_t(sd_count)));
              const std::vector<real_t> number = conc_to_number(cs, conc);
              std::vector<n_t> n(ijk.size());
              for (size_t s = 0; s < ijk.size(); ++s) n[s] = n_t(number[ijk[s]] / sd_count + real_t(.5));   // init_n.ipp:147-162
              finalize_new(cs, ijk, n, std::vector<real_t>(ijk.size(), radius * radius * radius), species.first.kappa);
            }
        }

        // ---- time stepping ---------------------------------------------------------------------------------------
        void adjust_timesteps(real_t dt_)   // impl_adjust_timesteps.ipp:13-22
        {
          if (dt_ > 0 && !oi.variable_dt_switch) throw std::runtime_error("libcloudph++: opts.dt specified, but opts_init.variable_dt_switch is false.");
          sstp_cond = dt_ > 0 && oi.sstp_cond > 1 ? int(std::ceil(oi.sstp_cond * dt_ / oi.dt)) : oi.sstp_cond;
          sstp_coal = dt_ > 0 && oi.sstp_coal > 1 ? int(std::ceil(oi.sstp_coal * dt_ / oi.dt)) : oi.sstp_coal;
          sstp_cond_act = dt_ > 0 && oi.sstp_cond_act > 1 ? int(std::ceil(oi.sstp_cond_act * dt_ / oi.dt)) : oi.sstp_cond_act;
          dt = dt_ > 0 ? dt_ : oi.dt;
        }

        // defer_wait: the caller (step_sync) runs step_cond right away and waits for the uploads at its end, so the Courant
        // fields - needed only by step_async - travel while the condensation kernel runs
        // argument and call-order checks of sync_in (particles_step.ipp:32-110), the Courant maps, the Euler fall-back of pred_corr
        bool sync_in_checks(arrinfo_t<real_t> &th, arrinfo_t<real_t> &rv, const arrinfo_t<real_t> &rhod, const arrinfo_t<real_t> &cx,
                            const arrinfo_t<real_t> &cy, const arrinfo_t<real_t> &cz, const arrinfo_t<real_t> &diss_rate, size_t n_ambient_chem)
        {
          if (!init_called) throw std::runtime_error("libcloudph++: please call init() before calling step_sync()");
          if (should_now_run_async) throw std::runtime_error("libcloudph++: please call step_async() before calling step_sync() again");
          if (th.is_null() || rv.is_null()) throw std::runtime_error("libcloudph++: passing th and rv is mandatory");
          check_courants(cx, cy, cz);
          if (n_ambient_chem != 0) throw std::runtime_error("libcloudph++: chemistry was switched off and ambient_chem is not empty");
          if (!diss_rate.is_null())
            throw std::runtime_error("libcloudph++: turbulent advection, coalescence and condesation are switched off and diss_rate is not empty");
          if (m_cx.l2e.empty()) init_courant_maps(cx, cy, cz);
          // |C_x| > 2 would leave the 2-cell halo of the predictor-corrector scheme: fall back to Euler for this step
          if (oi.adve_scheme == as_t::pred_corr && !cx.is_null())
          {
            real_t cmin = std::numeric_limits<real_t>::max(), cmax = -cmin;
            const host_view h_cx(cx, m_cx);
            for (const long l : m_cx.l2e) { cmin = std::min(cmin, h_cx.data[l]); cmax = std::max(cmax, h_cx.data[l]); }
            if (!(cmin >= real_t(-2.)) || !(cmax <= real_t(2.))) adve_scheme = as_t::euler;
          }
          return true;
        }

        // defer_wait: the caller (step_sync) runs step_cond right away and waits for the uploads at its end, so the Courant
        // fields - needed only by step_async - travel while the condensation kernel runs
        void sync_in(arrinfo_t<real_t> &th, arrinfo_t<real_t> &rv, const arrinfo_t<real_t> &rhod, const arrinfo_t<real_t> &cx,
                     const arrinfo_t<real_t> &cy, const arrinfo_t<real_t> &cz, const arrinfo_t<real_t> &diss_rate, size_t n_ambient_chem,
                     bool defer_wait = false, bool checked = false)
        {
          if (!checked) sync_in_checks(th, rv, rhod, cx, cy, cz, diss_rate, n_ambient_chem);
          var_rho = !rhod.is_null();
          sync_in_field(th, m_th, LCX_F_TH);
          sync_in_field(rv, m_rv, LCX_F_RV);
          sync_in_field(rhod, m_rhod, LCX_F_RHOD);
          sync_in_field(cx, m_cx, LCX_F_COURANT_X);
          sync_in_field(cy, m_cy, LCX_F_COURANT_Y);
          sync_in_field(cz, m_cz, LCX_F_COURANT_Z);
          transfers_open = defer_wait;
          if (!defer_wait) finish_transfers();          // the caller may overwrite its arrays as soon as this returns
          should_now_run_cond = true;
        }

        // ---- chunked step_sync ------------------------------------------------------------------------------------------
        // Condensation is cell-local, so step_sync need not be upload -> compute -> read back: the grid is cut into chunks of
        // whole kernel runs, chunk k + 1 travels to the device and chunk k's th / rv travel back while a chunk computes
        // (lcx_set_cell_window).  Same kernels on the same cells in the same order inside each run: bit-identical results.
        // Chosen by step_sync when the per-cell path with one sub-step runs on contiguous host arrays.
        // read at every call (two getenv per step) so that tests can switch it inside one process
        static long env_long(const char *name, long dflt) { const char *s = std::getenv(name); return s && *s ? std::atol(s) : dflt; }
        // piece [c0, c1) of a field's runs
        template <class F> static void for_run_pieces(const map_t &m, long c0, long c1, F f)
        {
          for (const run_t &r : m.runs)
          {
            const long lo = std::max(c0, r.dst), hi = std::min(c1, r.dst + r.len);
            if (lo < hi) f(lo, r.src + (lo - r.dst), hi - lo);
          }
        }
        bool step_sync_chunked(const opts_t<real_t> &opts, arrinfo_t<real_t> &th, arrinfo_t<real_t> &rv, const arrinfo_t<real_t> &rhod,
                               const arrinfo_t<real_t> &cx, const arrinfo_t<real_t> &cy, const arrinfo_t<real_t> &cz)
        {
          const long K = env_long("LCX_SYNC_CHUNKS", 8);
          if (K < 2 || !opts.cond || opts.turb_cond || oi.exact_sstp_cond || allow_sstp_cond || oi.const_p) return false;
          if (opts.chem_dsl || opts.chem_dsc || opts.chem_rct || long(n_cell) < env_long("LCX_SYNC_CHUNK_MIN_CELLS", 1l << 16)) return false;
          if (opts.dt > 0 || th.is_null() || rv.is_null() || !m_th.direct() || !m_rv.direct() || (!rhod.is_null() && !m_rhod.direct())) return false;
          if (on_device(th.data) || on_device(rv.data)) return false;      // device-resident fields: nothing to hide
          int64_t granule = 0;
          chk(L::cond_granule(e, &granule));
          if (granule <= 0) return false;
          // chunk boundaries, multiples of the granule: K equal chunks, except that the first and the last one are cut in four and two
          // (1/4, 1/2 of a chunk ... 1/2, 1/4): the upload of the first chunk and the read-back of the last are the part of the
          // traffic that no kernel hides
          const long n = long(n_cell);
          long chunk = (n + K - 1) / K;
          chunk = (chunk + granule - 1) / granule * granule;
          if (chunk >= n) return false;
          std::vector<long> edge(1, 0);
          auto cut = [&](long at) { at = std::min(n, (at + granule - 1) / granule * granule); if (at > edge.back()) edge.push_back(at); };
          const bool graded = env_long("LCX_SYNC_GRADED", 1) != 0 && chunk >= 8 * granule;
          for (long c0 = 0; c0 < n; c0 += chunk)
          {
            const long c1 = std::min(c0 + chunk, n);
            if (graded && c0 == 0) { cut(chunk / 4); cut(chunk / 4 + chunk / 2); }
            if (graded && c1 == n) { cut(c1 - (c1 - c0) / 4 - (c1 - c0) / 2); cut(c1 - (c1 - c0) / 4); }
            cut(c1);
          }

          // what sync_in does, except that th / rv / rhod go chunk by chunk and the Courant fields last
          var_rho = !rhod.is_null();
          should_now_run_cond = false;
          adjust_timesteps(opts.dt);
          auto upload = [&](long c0, long c1) {
            for_run_pieces(m_th, c0, c1, [&](long dst, long src, long len) { chk(L::cells_set_part(e, LCX_F_TH, dst, th.data + src, len)); });
            for_run_pieces(m_rv, c0, c1, [&](long dst, long src, long len) { chk(L::cells_set_part(e, LCX_F_RV, dst, rv.data + src, len)); });
            if (var_rho) for_run_pieces(m_rhod, c0, c1, [&](long dst, long src, long len) { chk(L::cells_set_part(e, LCX_F_RHOD, dst, rhod.data + src, len)); });
          };
          upload(edge[0], edge[1]);          // before anything is queued on the engine's stream
          chk(L::hskpng_mfp(e));             // from the T, p left by the previous Tpr (particles_step.ipp:189-194)
          try
          {
            for (size_t k = 0; k + 1 < edge.size(); ++k)
            {
              const long c0 = edge[k], c1 = edge[k + 1];
              chk(L::set_cell_window(e, c0, c1));
              chk(L::hskpng_Tpr(e));
              chk(L::cond(e, dt, opts.RH_max, 0, 1));
              if (k + 2 < edge.size()) upload(c1, edge[k + 2]);
              else
              {
                sync_in_field(cx, m_cx, LCX_F_COURANT_X);
                sync_in_field(cy, m_cy, LCX_F_COURANT_Y);
                sync_in_field(cz, m_cz, LCX_F_COURANT_Z);
              }
              for_run_pieces(m_th, c0, c1, [&](long cell, long host, long len) { chk(L::cells_get_part(e, LCX_F_TH, cell, th.data + host, len)); });
              for_run_pieces(m_rv, c0, c1, [&](long cell, long host, long len) { chk(L::cells_get_part(e, LCX_F_RV, cell, rv.data + host, len)); });
            }
          }
          catch (...) { L::set_cell_window(e, 0, 0); throw; }
          chk(L::set_cell_window(e, 0, 0));
          finish_transfers();
          transfers_open = false;
          should_now_run_async = true;
          return true;
        }

        // device-resident variant of sync_in + step_cond: the fields of the previous step stay where they are
        void step_resident_cond(const opts_t<real_t> &opts)
        {
          if (!init_called) throw std::runtime_error("libcloudph++: please call init() before calling step_sync()");
          if (should_now_run_async) throw std::runtime_error("libcloudph++: please call step_async() before calling step_sync() again");
          arrinfo_t<real_t> none_th, none_rv;
          var_rho = false;
          should_now_run_cond = true;
          step_cond(opts, none_th, none_rv);
        }

        void step_cond(const opts_t<real_t> &opts, arrinfo_t<real_t> &th, arrinfo_t<real_t> &rv)
        {
          if (!should_now_run_cond) throw std::runtime_error("libcloudph++: please call sync_in() before calling step_cond()");
          if (opts.turb_cond) throw std::runtime_error("libcloudph++: turb_cond_swtich=False, but turb_cond==True");
          should_now_run_cond = false;
          adjust_timesteps(opts.dt);
          if (opts.cond)
          {
            chk(L::hskpng_mfp(e));          // from the T, p left by the previous Tpr, as the reference does (particles_step.ipp:189-194)
            if (oi.exact_sstp_cond && (sstp_cond > 1 || sstp_cond_act > 1))     // per-particle sub-stepping: particles_step.ipp:199-236
            {
              if (oi.adaptive_sstp_cond)
                chk(L::cond_perparticle_adaptive(e, dt, opts.RH_max, sstp_cond, sstp_cond_act, oi.sstp_cond_adapt_drw2_eps, oi.sstp_cond_adapt_drw2_max));
              else
                chk(L::cond_perparticle(e, dt, opts.RH_max, sstp_cond, oi.sstp_cond_mix));
            }
            else
              for (int step = 0; step < sstp_cond; ++step)
              {
                chk(L::sstp_percell_step(e, step, sstp_cond, var_rho));
                chk(L::hskpng_Tpr(e));
                chk(L::cond(e, dt / sstp_cond, opts.RH_max, step, sstp_cond));      // includes update_th_rv
              }
            chk(L::sstp_save(e));
            sync_out_field(LCX_F_TH, m_th, th);
            sync_out_field(LCX_F_RV, m_rv, rv);
            finish_transfers();
            transfers_open = false;
          }
          if (transfers_open) { finish_transfers(); transfers_open = false; }     // deferred uploads of step_sync
          if (opts.chem_dsl || opts.chem_dsc || opts.chem_rct)
            throw std::runtime_error("libcloudph++: all chemistry was switched off in opts_init");
          should_now_run_async = true;
        }

        // everything of step_async up to (not including) migration / post_copy: particles_step.ipp:339-482
        void step_async_local(const opts_t<real_t> &opts)
        {
          if (!should_now_run_async) throw std::runtime_error("libcloudph++: please call step_sync() before calling step_async() again");
          should_now_run_async = false;
          if (opts.chem_dsl || opts.chem_dsc || opts.chem_rct) throw std::runtime_error("libcloudph++: all chemistry was switched off in opts_init");
          if (opts.coal && !oi.coal_switch) throw std::runtime_error("libcloudph++: coalescence was switched off in opts_init");
          if (opts.sedi && !oi.sedi_switch) throw std::runtime_error("libcloudph++: sedimentation was switched off in opts_init");
          if (opts.subs && !oi.subs_switch) throw std::runtime_error("libcloudph++: subsidence was switched off in opts_init");
          if (opts.turb_adve) throw std::runtime_error("libcloudph++: turb_adve_switch=False, but turb_adve==True");
          if (opts.src) throw std::runtime_error("libcloudph++: aerosol source was switched off in opts_init");
          if (opts.rlx) throw std::runtime_error("libcloudph++: aerosol relaxation was switched off in opts_init");
          adjust_timesteps(opts.dt);

          chk(L::hskpng_Tpr(e));
          if (opts.sedi || opts.coal || opts.cond) chk(L::hskpng_vterm(e, 0));

          if (opts.coal)
          {
            for (int step = 0; step < sstp_coal; ++step)
            {
              lcx_rng r;
              std::memset(&r, 0, sizeof(r));
              if (!injected.empty())
              {
                int64_t n_part = 0;
                chk(L::n_part(e, &n_part));
                un_host.swap(injected.front().first); u01_host.swap(injected.front().second);
                injected.pop_front();
                if (un_host.size() < size_t(n_part) || u01_host.size() < size_t(n_part))
                  throw std::runtime_error("libcloudph++ (B200 engine): injected random stream shorter than the number of super-droplets");
                if (!dense_sid) throw std::runtime_error("libcloudph++ (B200 engine): injected random streams need dense storage indices (lgrngn_b200_set_dense_sid)");
                r.mode = LCX_RNG_INJECT; r.un = un_host.data(); r.u01 = u01_host.data();
                ++philox_call;                 // a caller replaying the Philox stream on the host stays in step with the call counter
              }
              else if (rng_mode == LGRNGN_B200_RNG_MT19937)
              {
                // same draw order as the reference: un[n_part] for the shuffle, then u01[n_part] (hskpng_sort.ipp:33, coal.ipp:373)
                int64_t n_part = 0;
                chk(L::n_part(e, &n_part));
                un_host.resize(size_t(n_part)); u01_host.resize(size_t(n_part));
                std::uniform_int_distribution<unsigned int> dist_un(0, std::numeric_limits<unsigned int>::max());
                for (auto &v : un_host) v = uint32_t(real_t(dist_un(engine)));
                for (auto &v : u01_host) v = draw_u01();
                r.mode = LCX_RNG_INJECT; r.un = un_host.data(); r.u01 = u01_host.data();
              }
              else
              {
                // counter-based stream: key = (rng_seed, slab rank), counter = (global cell or storage index, block, call)
                r.mode = LCX_RNG_PHILOX; r.seed = uint64_t(uint32_t(oi.rng_seed)); r.call = philox_call++;
                r.cell_base = uint32_t(n_cell_bfr + philox_cell_base); r.stream = uint32_t(slab_rank);
              }
              chk(L::coal(e, dt / sstp_coal, &r));
              if (step + 1 != sstp_coal) chk(L::hskpng_vterm(e, 1));
            }
            if (pure_const_multi)
            {
              int flag = 0;
              chk(L::coal_flag(e, &flag));
              if (flag) ++sstp_coal;
            }
            chk(L::hskpng_rc2(e));                       // collisions changed rd3 / kappa (particles_step.ipp:402-403)
          }

          if (n_dims > 0)
          {
            // rank-local Courant arrays cannot hold the neighbours' columns of the predictor-corrector halo: exchanged between the
            // slabs' engines every step (particles_step.ipp:111-112 -> xchng_courants.ipp:15-153; here over peer memory, no MPI)
            if (process_distributed && halo_size > 0 && opts.adve) { chk(L::halo_put(e)); chk(L::halo_take(e)); }
            lcx_transport_opts t;
            t.adve = opts.adve; t.sedi = opts.sedi; t.subs = opts.subs; t.adve_scheme = int(adve_scheme); t.dt = dt;
            chk(L::transport(e, &t));
          }
          adve_scheme = oi.adve_scheme;
        }

        void post_copy(const opts_t<real_t> &opts) { chk(L::post_copy(e, opts.rcyc, 0)); }

        // x-slab migration through the neighbours' inboxes (include/lcx_b200.h); rgt / lft = neighbour engines of this process,
        // nullptr when the neighbours live in other processes
        int64_t n_sent[2] = {0, 0}, n_received[2] = {0, 0};
        void migr_put() { chk(L::migr_put(e, &n_sent[0], &n_sent[1])); }
        void migr_take(typename L::engine *rgt, typename L::engine *lft) { chk(L::migr_take(e, rgt, lft, &n_received[0], &n_received[1])); }

        bool process_distributed = false;      // one slab of a run spread over several processes (particles_b200.h)

        void step_async(const opts_t<real_t> &opts)
        {
          step_async_local(opts);
          if (process_distributed)
          {
            if (opts.rcyc) throw std::runtime_error("libcloudph++: Particle recycling can't be used in distributed-memory runs");
            // like the reference's MPI build, the exchange is part of step_async (particles_step.ipp:484-490 -> mpi_exchange)
            if (opts.adve) { migr_put(); migr_take(nullptr, nullptr); }
            post_copy(opts);
          }
          else if (!spawned) post_copy(opts);
        }

        // ---- diagnostics -----------------------------------------------------------------------------------------
        void diag_field(int field) { chk(L::hskpng_Tpr(e)); chk(L::diag_cell_field(e, field)); }

        real_t *outbuf()
        {
          outbuf_host.resize(n_cell);
          chk(L::outbuf(e, outbuf_host.data(), int64_t(n_cell)));
          return outbuf_host.data();
        }

        std::vector<real_t> get_attr(const std::string &name)
        {
          int a;
          if (name == "rw2") a = LCX_A_RW2; else if (name == "rd3") a = LCX_A_RD3; else if (name == "kappa") a = LCX_A_KPA;
          else if (name == "x") a = LCX_A_X; else if (name == "y") a = LCX_A_Y; else if (name == "z") a = LCX_A_Z;
          else if (name == "n") a = LCX_A_N; else if (name == "vt") a = LCX_A_VT; else if (name == "ijk") a = LCX_A_IJK;
          else if (name == "rd2_insol" || name == "T_freeze" || name == "ice_a" || name == "ice_c" || name == "ice_rho")
            throw std::runtime_error("Requested ice attribute '" + name + "' but ice_switch is off.");
          else throw std::runtime_error("Unknown attribute name passed to get_attr.");
          if ((a == LCX_A_X && !oi.nx) || (a == LCX_A_Y && !oi.ny) || (a == LCX_A_Z && !oi.nz)) return std::vector<real_t>();
          int64_t n_part = 0;
          chk(L::n_part(e, &n_part));
          std::vector<real_t> out(size_t(n_part), real_t(0));
          int64_t got = 0;
          chk(L::get_attr(e, a, out.data(), n_part, &got));
          return out;
        }

        std::map<common::output_t, real_t> diag_puddle()
        {
          double raw[14];
          chk(L::puddle(e, raw));
          std::map<common::output_t, real_t> res;
          for (int q = 0; q < 14; ++q) res[static_cast<common::output_t>(q)] = real_t(raw[q]);
          return res;
        }
      };

      // ======================================================================================================
      // particles_proto_t over one or several slabs
      // ======================================================================================================
      template <class real_t>
      struct particles_impl : particles_proto_t<real_t>
      {
        typedef particles_proto_t<real_t> parent_t;
        typedef typename parent_t::chem_map_t chem_map_t;
        typedef typename parent_t::chem_cmap_t chem_cmap_t;
        typedef lcx_api<real_t> L;
        static void chk(int rc) { slab<real_t>::chk(rc); }

        std::vector<std::unique_ptr<slab<real_t>>> slabs;
        std::unique_ptr<slab_workers> workers;     // multi_CUDA with more than one slab: one host thread per slab
        opts_init_t<real_t> glob;
        bool multi;
        bool connected = false;
        std::vector<real_t> gathered;

        template <class F>
        void for_slabs(F f)
        {
          if (workers) workers->run([&](int d) { f(*slabs[size_t(d)]); });
          else for (auto &s : slabs) f(*s);
        }

        // single device (optionally one rank of a process-distributed run, see particles_b200.h)
        explicit particles_impl(const opts_init_t<real_t> &o) : glob(o), multi(false)
        {
          std::pair<int, int> bc;
          const lgrngn_b200_distmem dm = g_distmem;
          g_distmem = lgrngn_b200_distmem{0, 1, -1., -1., 0};      // consumed: later particle systems of this process are ordinary ones
          if (dm.size > 1)
          {
            if (o.nx == 0) throw std::runtime_error("libcloudph++: distributed memory doesn't work for 0D setup.");
            if (!o.open_side_walls) bc = std::make_pair(int(distmem), int(distmem));
            else bc = std::make_pair(dm.rank == 0 ? int(open) : int(distmem), dm.rank == dm.size - 1 ? int(open) : int(distmem));
          }
          else bc = o.open_side_walls ? std::make_pair(int(open), int(open)) : std::make_pair(int(sharedmem), int(sharedmem));
          slabs.emplace_back(new slab<real_t>(o, bc, o.nx));   // Eulerian arrays are rank-local
          if (dm.size > 1)
          {
            slab<real_t> &s0 = *slabs[0];
            s0.spawned = true;
            s0.process_distributed = true;
            s0.lft_x1 = real_t(dm.lft_x1);
            s0.rgt_x0 = real_t(dm.rgt_x0);
            s0.slab_rank = dm.rank;
            // the Eulerian arrays are rank-local; the global cell offset only decorrelates the slabs' random streams
            s0.philox_cell_base = size_t(dm.rank) * s0.n_cell;
          }
          this->opts_init = &slabs[0]->oi;
        }

        // several devices in this process: particles_multi_gpu_impl.ipp:35-207
        particles_impl(const opts_init_t<real_t> &o, int) : glob(o), multi(true)
        {
          if (glob.nx == 0) throw std::runtime_error("libcloudph++: multi_CUDA doesn't work for 0D setup.");
          if (!(glob.x1 > glob.x0 && glob.x1 <= glob.nx * glob.dx)) throw std::runtime_error("libcloudph++: !(x1 > x0 & x1 <= min(1,nx)*dx)");
          // LCX_SLABS_ON_ONE_DEVICE=1 places every slab on device 0: lets the decomposition be tested on a single GPU
          const char *fold = std::getenv("LCX_SLABS_ON_ONE_DEVICE");
          const bool one_device = fold && std::string(fold) == "1";
          int dev_count = L::device_count();
          if (glob.dev_count > 0)
          {
            if (dev_count < glob.dev_count && !(one_device && dev_count > 0))
            { std::ostringstream s; s << "number of available GPUs (" << dev_count << ") smaller than number of GPUs defined in opts_init (" << glob.dev_count << ")"; throw std::runtime_error(s.str()); }
            dev_count = glob.dev_count;
          }
          if (dev_count == 0) throw std::runtime_error("libcloudph++: no CUDA device is available");
          if (dev_count > glob.nx)
          { std::ostringstream s; s << "Number of CUDA devices (" << dev_count << ") used is greater than nx (" << glob.nx << ")"; throw std::runtime_error(s.str()); }
          glob.dev_count = dev_count;
          if (glob.dev_id >= 0)
          {
            std::cout << "Libcloudph++ warning: opts_init.dev_id is not compatible with the multi_CUDA backend, ignoring it's value." << std::endl;
            glob.dev_id = -1;
          }
          for (int d = 0; d < dev_count; ++d)
          {
            opts_init_t<real_t> o_d(glob);
            int n_x_bfr = 0;
            if (dev_count > 1) n_x_bfr = distmem_opts(o_d, d, dev_count);
            o_d.dev_id = one_device ? 0 : d;
            std::pair<int, int> bc = o_d.open_side_walls ? std::make_pair(int(open), int(open)) : std::make_pair(int(sharedmem), int(sharedmem));
            if (dev_count > 1)
            {
              if (d == 0) bc.second = distmem;
              else if (d == dev_count - 1) bc.first = distmem;
              else bc = std::make_pair(int(distmem), int(distmem));
              if (!o_d.open_side_walls) { if (d == 0) bc.first = distmem; else if (d == dev_count - 1) bc.second = distmem; }
            }
            slabs.emplace_back(new slab<real_t>(o_d, bc, glob.nx));
            slab<real_t> &s = *slabs.back();
            s.n_x_bfr = n_x_bfr;
            s.n_cell_bfr = size_t(n_x_bfr) * m1(o_d.ny) * m1(o_d.nz);
            s.spawned = dev_count > 1;
            s.slab_rank = d;
            s.oi.dev_count = dev_count;
          }
          if (dev_count > 1) workers.reset(new slab_workers(dev_count));
          for (int d = 0; d < dev_count && dev_count > 1; ++d)
          {
            const int lft = d > 0 ? d - 1 : dev_count - 1, rgt = d < dev_count - 1 ? d + 1 : 0;
            slabs[size_t(d)]->lft_x1 = slabs[size_t(lft)]->oi.x1;
            slabs[size_t(d)]->rgt_x0 = slabs[size_t(rgt)]->oi.x0;
          }
          this->opts_init = &glob;
        }

        slab<real_t> &one() { return *slabs[0]; }

        ~particles_impl() { workers.reset(); }

        // ---- API ---------------------------------------------------------------------------------------------------
        void init(const arrinfo_t<real_t> th, const arrinfo_t<real_t> rv, const arrinfo_t<real_t> rhod, const arrinfo_t<real_t> p,
                  const arrinfo_t<real_t> cx, const arrinfo_t<real_t> cy, const arrinfo_t<real_t> cz, const chem_cmap_t ambient_chem) override
        {
          const size_t n_chem = ambient_chem.size();
          for_slabs([&](slab<real_t> &s) { s.init(th, rv, rhod, p, cx, cy, cz, n_chem); });
          // every slab delivers its leavers straight into its neighbours' inboxes: particles_multi_gpu_impl.ipp:92-116 (peer access)
          const int G = int(slabs.size());
          if (multi && G > 1)
          {
            for (int d = 0; d < G; ++d)
            {
              const int lft = d > 0 ? d - 1 : G - 1, rgt = d < G - 1 ? d + 1 : 0;
              if (slabs[size_t(d)]->bcond.first == distmem) chk(L::migr_connect(slabs[size_t(d)]->e, 0, slabs[size_t(lft)]->e));
              if (slabs[size_t(d)]->bcond.second == distmem) chk(L::migr_connect(slabs[size_t(d)]->e, 1, slabs[size_t(rgt)]->e));
            }
            connected = true;
          }
        }

        void step_sync(const opts_t<real_t> &opts, arrinfo_t<real_t> th, arrinfo_t<real_t> rv, const arrinfo_t<real_t> rhod,
                       const arrinfo_t<real_t> cx, const arrinfo_t<real_t> cy, const arrinfo_t<real_t> cz,
                       const arrinfo_t<real_t> diss_rate, chem_map_t ambient_chem) override
        {
          const size_t n_chem = ambient_chem.size();
          for_slabs([&](slab<real_t> &s) {
            arrinfo_t<real_t> th_ = th, rv_ = rv;
            if (s.sync_in_checks(th_, rv_, rhod, cx, cy, cz, diss_rate, n_chem) && s.step_sync_chunked(opts, th_, rv_, rhod, cx, cy, cz)) return;
            s.sync_in(th_, rv_, rhod, cx, cy, cz, diss_rate, n_chem, /*defer_wait=*/true, /*checked=*/true);
            s.step_cond(opts, th_, rv_);
          });
        }

        void sync_in(arrinfo_t<real_t> th, arrinfo_t<real_t> rv, const arrinfo_t<real_t> rhod, const arrinfo_t<real_t> cx,
                     const arrinfo_t<real_t> cy, const arrinfo_t<real_t> cz, const arrinfo_t<real_t> diss_rate, chem_map_t ambient_chem) override
        {
          const size_t n_chem = ambient_chem.size();
          for_slabs([&](slab<real_t> &s) { arrinfo_t<real_t> th_ = th, rv_ = rv; s.sync_in(th_, rv_, rhod, cx, cy, cz, diss_rate, n_chem); });
        }

        void step_cond(const opts_t<real_t> &opts, arrinfo_t<real_t> th, arrinfo_t<real_t> rv, chem_map_t) override
        {
          for_slabs([&](slab<real_t> &s) { arrinfo_t<real_t> th_ = th, rv_ = rv; s.step_cond(opts, th_, rv_); });
        }

        void step_async(const opts_t<real_t> &opts) override
        {
          if (multi && opts.rcyc)
            throw std::runtime_error("libcloudph++: Particle recycling can't be used in the multi_CUDA backend (it would consume whole memory quickly");
          const int G = int(slabs.size());
          if (multi && G > 1)
          {
            // phase 1: everything local, then each slab packs its leavers straight into its neighbours' inboxes (peer memory);
            // phase 2 (after all deliveries have been queued - run() returning is the barrier): each slab's stream waits for its
            // neighbours' delivery events, appends the arrivals (right neighbour's first) and re-groups.  The reference needs five
            // barriers and four peer copies per step for this (step_async_and_copy.ipp:28-206).
            for_slabs([&](slab<real_t> &s) { s.step_async_local(opts); if (opts.adve) s.migr_put(); });
            for_slabs([&](slab<real_t> &s) {
              if (opts.adve)
              {
                const int d = s.slab_rank, lft = d > 0 ? d - 1 : G - 1, rgt = d < G - 1 ? d + 1 : 0;
                s.migr_take(s.bcond.second == distmem ? slabs[size_t(rgt)]->e : nullptr, s.bcond.first == distmem ? slabs[size_t(lft)]->e : nullptr);
              }
              s.post_copy(opts);
            });
          }
          else
            for (auto &s : slabs) s->step_async(opts);
        }

        // one model step on the fields already resident in device memory (particles_b200.h)
        void step_resident(const opts_t<real_t> &opts)
        {
          for_slabs([&](slab<real_t> &s) { s.step_resident_cond(opts); });
          step_async(opts);
        }

        // diagnostics need the engines, which init() creates (the reference asserts here; a release build of it would crash)
        void ready() const
        {
          for (const auto &s : slabs)
            if (!s->e) throw std::runtime_error("libcloudph++: please call init() before asking for diagnostics");
        }

        // selectors
        void sel(int kind, int attr, real_t lo, real_t hi, bool cons) { ready(); for (auto &s : slabs) chk(L::moms_select(s->e, kind, attr, lo, hi, cons)); }
        void diag_all() override { sel(LCX_SEL_ALL, 0, 0, 0, false); }
        void diag_rw_ge_rc() override { sel(LCX_SEL_RW_GE_RC, 0, 0, 0, false); }
        void diag_RH_ge_Sc() override { sel(LCX_SEL_RH_GE_SC, 0, 0, 0, false); }
        void diag_dry_rng(const real_t &r0, const real_t &r1) override { sel(LCX_SEL_RANGE, LCX_A_RD3, std::pow(r0, 3), std::pow(r1, 3), false); }
        void diag_wet_rng(const real_t &r0, const real_t &r1) override { sel(LCX_SEL_RANGE, LCX_A_RW2, std::pow(r0, 2), std::pow(r1, 2), false); }
        void diag_kappa_rng(const real_t &k0, const real_t &k1) override { sel(LCX_SEL_RANGE, LCX_A_KPA, k0, k1, false); }
        void diag_water() override { sel(LCX_SEL_GT0, LCX_A_RW2, 0, 0, false); }
        void diag_dry_rng_cons(const real_t &r0, const real_t &r1) override { sel(LCX_SEL_RANGE, LCX_A_RD3, std::pow(r0, 3), std::pow(r1, 3), true); }
        void diag_wet_rng_cons(const real_t &r0, const real_t &r1) override { sel(LCX_SEL_RANGE, LCX_A_RW2, std::pow(r0, 2), std::pow(r1, 2), true); }
        void diag_kappa_rng_cons(const real_t &k0, const real_t &k1) override { sel(LCX_SEL_RANGE, LCX_A_KPA, k0, k1, true); }
        void diag_water_cons() override { sel(LCX_SEL_GT0, LCX_A_RW2, 0, 0, true); }
        void diag_ice() override { no_ice(); }
        void diag_ice_cons() override { no_ice(); }
        void diag_ice_a_rng(const real_t &, const real_t &) override { no_ice(); }
        void diag_ice_c_rng(const real_t &, const real_t &) override { no_ice(); }
        void diag_ice_a_rng_cons(const real_t &, const real_t &) override { no_ice(); }
        void diag_ice_c_rng_cons(const real_t &, const real_t &) override { no_ice(); }
        void diag_ice_a_mom(const int &) override { no_ice(); }
        void diag_ice_c_mom(const int &) override { no_ice(); }
        void diag_ice_mix_ratio() override { no_ice(); }
        void diag_precip_rate_ice_mass() override { no_ice(); }
        static void no_ice() { throw std::runtime_error("libcloudph++: ice is switched off in opts_init, but diag_ice was called"); }
        void diag_chem(const enum common::chem::chem_species_t &) override
        { throw std::runtime_error("libcloudph++: chemistry is switched off in opts_init, but diag_chem was called"); }

        // moments and fields
        void mom(int attr, real_t power) { ready(); for (auto &s : slabs) chk(L::moms_calc(s->e, attr, power, 1)); }
        void diag_dry_mom(const int &k) override { mom(LCX_A_RD3, k / 3.); }
        void diag_wet_mom(const int &k) override { mom(LCX_A_RW2, k / 2.); }
        void diag_kappa_mom(const int &k) override { mom(LCX_A_KPA, k); }
        void diag_sd_conc() override { ready(); for (auto &s : slabs) chk(L::diag_sd_conc(s->e)); }
        void diag_pressure() override { ready(); for (auto &s : slabs) s->diag_field(LCX_F_P); }
        void diag_temperature() override { ready(); for (auto &s : slabs) s->diag_field(LCX_F_T); }
        void diag_RH() override { ready(); for (auto &s : slabs) s->diag_field(LCX_F_RH); }
        void diag_precip_rate() override { ready(); for (auto &s : slabs) chk(L::diag_precip_rate(s->e)); }
        void diag_max_rw() override { ready(); for (auto &s : slabs) chk(L::diag_max_rw(s->e)); }
        void diag_wet_mass_dens(const real_t &rad, const real_t &sig0) override { ready(); for (auto &s : slabs) chk(L::diag_mass_dens(s->e, LCX_A_RW2, rad, sig0, 1. / 2.)); }
        void diag_vel_div() override { ready(); for (auto &s : slabs) chk(L::diag_vel_div(s->e, s->oi.dt)); }

        real_t *outbuf() override
        {
          ready();
          if (!multi) return one().outbuf();
          gathered.resize(size_t(m1(glob.nx)) * m1(glob.ny) * m1(glob.nz));
          for (auto &s : slabs)
          {
            const real_t *part = s->outbuf();
            std::copy(part, part + s->n_cell, gathered.begin() + long(s->n_cell_bfr));
          }
          return gathered.data();
        }

        std::vector<real_t> get_attr(const std::string &name) override
        {
          if (multi) throw std::runtime_error("get_attr doesnt work in multi_CUDA backend.");
          ready();
          return one().get_attr(name);
        }

        std::map<common::output_t, real_t> diag_puddle() override
        {
          ready();
          std::map<common::output_t, real_t> res;
          for (int q = 0; q < 14; ++q) res[static_cast<common::output_t>(q)] = 0;
          for (auto &s : slabs) for (const auto &kv : s->diag_puddle()) res[kv.first] += kv.second;
          return res;
        }
      };
    }


    // ---- single-precision callers ----------------------------------------------------------------------------------
    // factory<float> (src/lib.cpp:43) is served by the double-precision engine: options and fields are widened on the way in,
    // th / rv, outbuf, get_attr and diag_puddle narrowed on the way out.  Arithmetic is therefore done in double (compute
    // type >= the caller's); super-droplet attributes are not bit-comparable with a float build of the reference (whose
    // random draws and roundings differ), fields and moments agree to single-precision accuracy.
    namespace b200
    {
      class particles_float : public particles_proto_t<float>
      {
        typedef particles_proto_t<float> base_t;
        std::unique_ptr<particles_proto_t<double>> dbl;
        opts_init_t<float> oi_f;
        std::vector<float> outbuf_f;
        int n_dims;

        struct widened_fun : common::unary_function<double>
        {
          std::shared_ptr<common::unary_function<float>> f;
          double funval(const double x) const { return double((*f)(float(x))); }
        };

        struct staged
        {
          std::vector<double> buf;
          std::vector<ptrdiff_t> strides;
          arrinfo_t<double> info() { return buf.empty() ? arrinfo_t<double>() : arrinfo_t<double>(buf.data(), strides); }
        };

        // extents of a field: scalars nx x ny x nz, Courant components one more along their own direction
        void extents(int ext, long (&n)[3], int &nd) const
        {
          const int nn[3] = {oi_f.nx, oi_f.ny, oi_f.nz};
          nd = 0;
          for (int d = 0; d < 3; ++d) if (nn[d] > 0) n[nd++] = nn[d] + (ext == d ? 1 : 0);
          for (int d = nd; d < 3; ++d) n[d] = 1;
        }
        template <class F>
        void for_each_element(const arrinfo_t<float> &a, int ext, F f) const
        {
          long n[3]; int nd;
          extents(ext, n, nd);
          const ptrdiff_t s0 = nd > 0 ? (nd == 1 ? 1 : a.strides[0]) : 0, s1 = nd > 1 ? a.strides[1] : 0, s2 = nd > 2 ? a.strides[2] : 0;
          long q = 0;
          for (long i = 0; i < n[0]; ++i)
            for (long j = 0; j < n[1]; ++j)
              for (long k = 0; k < n[2]; ++k, ++q)
                f(q, a.data[i * s0 + j * s1 + k * s2]);
        }
        void widen(const arrinfo_t<float> &a, int ext, staged &st) const
        {
          st.buf.clear();
          if (a.is_null()) return;
          long n[3]; int nd;
          extents(ext, n, nd);
          st.buf.resize(size_t(n[0] * n[1] * n[2]));
          st.strides.assign({ptrdiff_t(n[1] * n[2]), ptrdiff_t(nd > 2 ? n[2] : 1), ptrdiff_t(1)});
          if (nd == 2) st.strides = {ptrdiff_t(n[1]), ptrdiff_t(1)};
          if (nd <= 1) st.strides = {ptrdiff_t(1)};
          for_each_element(a, ext, [&](long q, float &v) { st.buf[size_t(q)] = double(v); });
        }
        void narrow(const staged &st, arrinfo_t<float> &a) const
        {
          if (a.is_null() || st.buf.empty()) return;
          for_each_element(a, -1, [&](long q, float &v) { v = float(st.buf[size_t(q)]); });
        }
        static opts_t<double> widen(const opts_t<float> &o)
        {
          opts_t<double> r;
          r.adve = o.adve; r.sedi = o.sedi; r.subs = o.subs; r.cond = o.cond; r.coal = o.coal; r.src = o.src; r.rlx = o.rlx; r.rcyc = o.rcyc;
          r.turb_adve = o.turb_adve; r.turb_cond = o.turb_cond; r.turb_coal = o.turb_coal; r.ice_nucl = o.ice_nucl;
          r.chem_dsl = o.chem_dsl; r.chem_dsc = o.chem_dsc; r.chem_rct = o.chem_rct;
          r.RH_max = o.RH_max; r.dt = o.dt;
          return r;
        }
        static opts_init_t<double> widen(const opts_init_t<float> &o)
        {
          opts_init_t<double> r;
          for (const auto &dd : o.dry_distros)
          {
            auto w = std::make_shared<widened_fun>();
            w->f = dd.second;
            r.dry_distros.emplace(kappa_rd_insol_t<double>(dd.first.kappa, dd.first.rd_insol), w);
          }
          for (const auto &ds : o.dry_sizes)
            for (const auto &sz : ds.second)
              r.dry_sizes[kappa_rd_insol_t<double>(ds.first.kappa, ds.first.rd_insol)][double(sz.first)] = std::make_pair(double(sz.second.first), sz.second.second);
          r.nx = o.nx; r.ny = o.ny; r.nz = o.nz; r.dx = o.dx; r.dy = o.dy; r.dz = o.dz; r.dt = o.dt;
          r.sstp_cond = o.sstp_cond; r.sstp_coal = o.sstp_coal; r.sstp_cond_act = o.sstp_cond_act; r.sstp_chem = o.sstp_chem;
          r.x0 = o.x0; r.y0 = o.y0; r.z0 = o.z0; r.x1 = o.x1; r.y1 = o.y1; r.z1 = o.z1;
          r.sd_conc = o.sd_conc; r.sd_conc_large_tail = o.sd_conc_large_tail; r.aerosol_independent_of_rhod = o.aerosol_independent_of_rhod;
          r.variable_dt_switch = o.variable_dt_switch; r.sd_const_multi = o.sd_const_multi; r.n_sd_max = o.n_sd_max;
          r.kernel = o.kernel; r.terminal_velocity = o.terminal_velocity; r.adve_scheme = o.adve_scheme; r.RH_formula = o.RH_formula;
          r.kernel_parameters.assign(o.kernel_parameters.begin(), o.kernel_parameters.end());
          r.chem_switch = o.chem_switch; r.coal_switch = o.coal_switch; r.sedi_switch = o.sedi_switch; r.subs_switch = o.subs_switch;
          r.rlx_switch = o.rlx_switch; r.turb_adve_switch = o.turb_adve_switch; r.turb_cond_switch = o.turb_cond_switch;
          r.turb_coal_switch = o.turb_coal_switch; r.ice_switch = o.ice_switch; r.exact_sstp_cond = o.exact_sstp_cond;
          r.sstp_cond_mix = o.sstp_cond_mix; r.adaptive_sstp_cond = o.adaptive_sstp_cond; r.time_dep_ice_nucl = o.time_dep_ice_nucl;
          r.sstp_cond_adapt_drw2_eps = o.sstp_cond_adapt_drw2_eps; r.sstp_cond_adapt_drw2_max = o.sstp_cond_adapt_drw2_max;
          r.chem_rho = o.chem_rho; r.diag_incloud_time = o.diag_incloud_time;
          r.RH_max = o.RH_max; r.rng_seed = o.rng_seed; r.rng_seed_init = o.rng_seed_init; r.rng_seed_init_switch = o.rng_seed_init_switch;
          r.dev_count = o.dev_count; r.dev_id = o.dev_id;
          r.w_LS.assign(o.w_LS.begin(), o.w_LS.end()); r.SGS_mix_len.assign(o.SGS_mix_len.begin(), o.SGS_mix_len.end());
          r.aerosol_conc_factor.assign(o.aerosol_conc_factor.begin(), o.aerosol_conc_factor.end());
          r.rd_min = o.rd_min; r.rd_max = o.rd_max;
          r.no_ccn_at_init = o.no_ccn_at_init; r.open_side_walls = o.open_side_walls; r.periodic_topbot_walls = o.periodic_topbot_walls;
          r.rc2_T = o.rc2_T; r.src_type = o.src_type;
          r.rlx_bins = o.rlx_bins; r.rlx_sd_per_bin = o.rlx_sd_per_bin; r.supstp_rlx = o.supstp_rlx; r.rlx_timescale = o.rlx_timescale;
          r.th_dry = o.th_dry; r.const_p = o.const_p;
          return r;
        }

        staged s_th, s_rv, s_rhod, s_p, s_cx, s_cy, s_cz;

        public:
        particles_float(const backend_t backend, const opts_init_t<float> &o) : oi_f(o)
        {
          n_dims = (o.nx > 0) + (o.ny > 0) + (o.nz > 0);
          const opts_init_t<double> od = widen(o);
          dbl.reset(backend == multi_CUDA ? new particles_impl<double>(od, 0) : new particles_impl<double>(od));
          oi_f.n_sd_max = dbl->opts_init->n_sd_max;
          this->opts_init = &oi_f;
        }

        void init(const arrinfo_t<float> th, const arrinfo_t<float> rv, const arrinfo_t<float> rhod, const arrinfo_t<float> p,
                  const arrinfo_t<float> cx, const arrinfo_t<float> cy, const arrinfo_t<float> cz, const base_t::chem_cmap_t chem) override
        {
          if (!chem.empty()) throw std::runtime_error("libcloudph++: chemistry was switched off and ambient_chem is not empty");
          widen(th, -1, s_th); widen(rv, -1, s_rv); widen(rhod, -1, s_rhod); widen(p, -1, s_p);
          widen(cx, 0, s_cx); widen(cy, 1, s_cy); widen(cz, 2, s_cz);
          dbl->init(s_th.info(), s_rv.info(), s_rhod.info(), s_p.info(), s_cx.info(), s_cy.info(), s_cz.info());
        }
        void sync_in(arrinfo_t<float> th, arrinfo_t<float> rv, const arrinfo_t<float> rhod, const arrinfo_t<float> cx, const arrinfo_t<float> cy,
                     const arrinfo_t<float> cz, const arrinfo_t<float> diss_rate, base_t::chem_map_t chem) override
        {
          if (!chem.empty()) throw std::runtime_error("libcloudph++: chemistry was switched off and ambient_chem is not empty");
          if (!diss_rate.is_null())
            throw std::runtime_error("libcloudph++: turbulent advection, coalescence and condesation are switched off and diss_rate is not empty");
          widen(th, -1, s_th); widen(rv, -1, s_rv); widen(rhod, -1, s_rhod);
          widen(cx, 0, s_cx); widen(cy, 1, s_cy); widen(cz, 2, s_cz);
          dbl->sync_in(s_th.info(), s_rv.info(), s_rhod.info(), s_cx.info(), s_cy.info(), s_cz.info());
        }
        void step_cond(const opts_t<float> &o, arrinfo_t<float> th, arrinfo_t<float> rv, base_t::chem_map_t) override
        {
          dbl->step_cond(widen(o), s_th.info(), s_rv.info());
          narrow(s_th, th); narrow(s_rv, rv);
        }
        void step_sync(const opts_t<float> &o, arrinfo_t<float> th, arrinfo_t<float> rv, const arrinfo_t<float> rhod, const arrinfo_t<float> cx,
                       const arrinfo_t<float> cy, const arrinfo_t<float> cz, const arrinfo_t<float> diss_rate, base_t::chem_map_t chem) override
        {
          sync_in(th, rv, rhod, cx, cy, cz, diss_rate, chem);
          step_cond(o, th, rv, chem);
        }
        void step_async(const opts_t<float> &o) override { dbl->step_async(widen(o)); }

        void diag_all() override { dbl->diag_all(); }
        void diag_rw_ge_rc() override { dbl->diag_rw_ge_rc(); }
        void diag_RH_ge_Sc() override { dbl->diag_RH_ge_Sc(); }
        void diag_dry_rng(const float &a, const float &b) override { dbl->diag_dry_rng(a, b); }
        void diag_wet_rng(const float &a, const float &b) override { dbl->diag_wet_rng(a, b); }
        void diag_kappa_rng(const float &a, const float &b) override { dbl->diag_kappa_rng(a, b); }
        void diag_water() override { dbl->diag_water(); }
        void diag_dry_rng_cons(const float &a, const float &b) override { dbl->diag_dry_rng_cons(a, b); }
        void diag_wet_rng_cons(const float &a, const float &b) override { dbl->diag_wet_rng_cons(a, b); }
        void diag_kappa_rng_cons(const float &a, const float &b) override { dbl->diag_kappa_rng_cons(a, b); }
        void diag_water_cons() override { dbl->diag_water_cons(); }
        void diag_sd_conc() override { dbl->diag_sd_conc(); }
        void diag_pressure() override { dbl->diag_pressure(); }
        void diag_temperature() override { dbl->diag_temperature(); }
        void diag_RH() override { dbl->diag_RH(); }
        void diag_dry_mom(const int &k) override { dbl->diag_dry_mom(k); }
        void diag_wet_mom(const int &k) override { dbl->diag_wet_mom(k); }
        void diag_kappa_mom(const int &k) override { dbl->diag_kappa_mom(k); }
        void diag_wet_mass_dens(const float &a, const float &b) override { dbl->diag_wet_mass_dens(a, b); }
        void diag_precip_rate() override { dbl->diag_precip_rate(); }
        void diag_max_rw() override { dbl->diag_max_rw(); }
        void diag_vel_div() override { dbl->diag_vel_div(); }
        std::map<common::output_t, float> diag_puddle() override
        {
          std::map<common::output_t, float> r;
          for (const auto &kv : dbl->diag_puddle()) r[kv.first] = float(kv.second);
          return r;
        }
        std::vector<float> get_attr(const std::string &name) override
        {
          const std::vector<double> v = dbl->get_attr(name);
          return std::vector<float>(v.begin(), v.end());
        }
        float *outbuf() override
        {
          const double *src = dbl->outbuf();
          const size_t n = size_t(std::max(1, oi_f.nx)) * size_t(std::max(1, oi_f.ny)) * size_t(std::max(1, oi_f.nz));
          outbuf_f.resize(n);
          for (size_t q = 0; q < n; ++q) outbuf_f[q] = float(src[q]);
          return outbuf_f.data();
        }
      };
    }

    // ---- factory: src/lib.cpp:13-44 -----------------------------------------------------------------------------
    template <typename real_t>
    particles_proto_t<real_t> *factory(const backend_t backend, opts_init_t<real_t> opts_init)
    {
      switch (backend)
      {
        case multi_CUDA:
        case CUDA:
          // factory<float> (src/lib.cpp:43): the single-precision engine, like the reference's float instantiation.
          // LCX_FLOAT_VIA_DOUBLE=1 serves float callers from the double-precision engine instead (widening adapter).
          if constexpr (std::is_same<real_t, float>::value)
          {
            const char *v = std::getenv("LCX_FLOAT_VIA_DOUBLE");
            if (v && v[0] == '1') return new b200::particles_float(backend, opts_init);
          }
          if (backend == multi_CUDA) return new b200::particles_impl<real_t>(opts_init, 0);
          return new b200::particles_impl<real_t>(opts_init);
        case OpenMP:     throw std::runtime_error("libcloudph++: OpenMP backend was not compiled");
        case serial:     throw std::runtime_error("libcloudph++: serial backend was not compiled");
        default:         throw std::runtime_error("libcloudph++: unknown backend");
      }
    }

    template particles_proto_t<float> *factory(const backend_t, opts_init_t<float>);
    template particles_proto_t<double> *factory(const backend_t, opts_init_t<double>);
  }
}

// ---- C entry points for process-distributed runs and test control (particles_b200.h) ---------------------------------
namespace lg = libcloudphxx::lgrngn;

template <class F>
static int guarded_c(F f)
{
  try { f(); return 0; }
  catch (const std::exception &ex) { std::cerr << ex.what() << std::endl; return 1; }
}

template <class real_t>
static lg::b200::particles_impl<real_t> &impl_of_t(void *proto)
{
  auto *p = dynamic_cast<lg::b200::particles_impl<real_t> *>(static_cast<lg::particles_proto_t<real_t> *>(proto));
  if (!p) throw std::runtime_error("not a B200 particle system of this precision");
  return *p;
}

extern "C" {

void lgrngn_b200_set_rng_mode(int mode) { lg::b200::g_rng_mode = mode; }
int lgrngn_b200_get_rng_mode(void) { return lg::b200::g_rng_mode; }
void lgrngn_b200_set_distmem(const lgrngn_b200_distmem *d) { lg::b200::g_distmem = *d; }

void lgrngn_b200_set_dense_sid(int mode) { lg::b200::g_dense_sid = mode; }

static lg::b200::particles_impl<double> &impl_of(void *proto) { return impl_of_t<double>(proto); }
static lg::b200::slab<double> &slab_of(void *proto) { return impl_of(proto).one(); }

// ---- the same handles for a single-precision particle system (particles_proto_t<float>*) ----
int lgrngn_b200_n_slabs_f32(void *proto)
{
  try { return int(impl_of_t<float>(proto).slabs.size()); } catch (...) { return 0; }
}
void *lgrngn_b200_engine_of_slab_f32(void *proto, int d)      // an lcx_engine of liblcx_b200_f32.so: use the _f32 entry points on it
{
  try { auto &p = impl_of_t<float>(proto); return (d >= 0 && size_t(d) < p.slabs.size()) ? p.slabs[size_t(d)]->e : nullptr; } catch (...) { return nullptr; }
}
int lgrngn_b200_step_resident_f32(void *proto, int flags)
{
  return guarded_c([&] {
    lg::opts_t<float> o;
    o.adve = flags & 1; o.sedi = flags & 2; o.cond = flags & 4; o.coal = flags & 8;
    impl_of_t<float>(proto).step_resident(o);
  });
}
// multiplicities as 64-bit integers in storage order (get_attr("n") rounds them to real_t); real_bytes = sizeof the caller's real_t
int lgrngn_b200_get_n(void *proto, int real_bytes, unsigned long long *dst, long long cap, long long *n_out)
{
  return guarded_c([&] {
    int64_t got = 0;
    if (real_bytes == 4)
    {
      auto &s = impl_of_t<float>(proto).one();
      lg::b200::slab<float>::chk(lcx_get_attr_u64_f32(s.e, LCX_A_N, reinterpret_cast<uint64_t *>(dst), cap, &got));
    }
    else
      lg::b200::chk(lcx_get_attr_u64(slab_of(proto).e, LCX_A_N, reinterpret_cast<uint64_t *>(dst), cap, &got));
    *n_out = got;
  });
}

void *lgrngn_b200_engine(void *proto)
{
  try { return slab_of(proto).e; } catch (...) { return nullptr; }
}

int lgrngn_b200_n_slabs(void *proto)
{
  try { return int(impl_of(proto).slabs.size()); } catch (...) { return 0; }
}

void *lgrngn_b200_engine_of_slab(void *proto, int d)
{
  try { auto &p = impl_of(proto); return (d >= 0 && size_t(d) < p.slabs.size()) ? p.slabs[size_t(d)]->e : nullptr; } catch (...) { return nullptr; }
}

int lgrngn_b200_step_resident(void *proto, int flags)
{
  return guarded_c([&] {
    lg::opts_t<double> o;
    o.adve = flags & 1; o.sedi = flags & 2; o.cond = flags & 4; o.coal = flags & 8;
    impl_of(proto).step_resident(o);
  });
}

int lgrngn_b200_post_copy(void *proto, int rcyc)
{
  return guarded_c([&] { lg::opts_t<double> o; o.rcyc = rcyc != 0; slab_of(proto).post_copy(o); });
}

int lgrngn_b200_distmem_export(void *proto, void *blobs)
{
  return guarded_c([&] {
    auto &s = slab_of(proto);
    if (!s.e) throw std::runtime_error("libcloudph++ (B200 engine): call init() before exporting the migration inboxes");
    std::memset(blobs, 0, 2 * LCX_IPC_BLOB_BYTES);
    // inbox 0 receives the RIGHT neighbour's left-movers, inbox 1 the LEFT neighbour's right-movers
    if (s.bcond.second == lg::b200::distmem) lg::b200::chk(lcx_migr_ipc_export(s.e, 0, blobs));
    if (s.bcond.first == lg::b200::distmem) lg::b200::chk(lcx_migr_ipc_export(s.e, 1, static_cast<char *>(blobs) + LCX_IPC_BLOB_BYTES));
  });
}

int lgrngn_b200_distmem_connect(void *proto, const void *lft_blobs, const void *rgt_blobs)
{
  return guarded_c([&] {
    auto &s = slab_of(proto);
    if (!s.e) throw std::runtime_error("libcloudph++ (B200 engine): call init() before connecting the neighbours");
    if (s.bcond.first == lg::b200::distmem) lg::b200::chk(lcx_migr_ipc_connect(s.e, 0, lft_blobs));
    if (s.bcond.second == lg::b200::distmem) lg::b200::chk(lcx_migr_ipc_connect(s.e, 1, static_cast<const char *>(rgt_blobs) + LCX_IPC_BLOB_BYTES));
  });
}

int lgrngn_b200_migr_stats(void *proto, long long *out)
{
  return guarded_c([&] {
    auto &p = impl_of(proto);
    for (int q = 0; q < 4; ++q) out[q] = 0;
    for (auto &s : p.slabs) { out[0] += s->n_sent[0]; out[1] += s->n_sent[1]; out[2] += s->n_received[0]; out[3] += s->n_received[1]; }
  });
}

int lgrngn_b200_inject_rng(void *proto, const unsigned int *un, const double *u01, long long n)
{
  return guarded_c([&] {
    auto &s = slab_of(proto);
    s.injected.emplace_back(std::vector<uint32_t>(un, un + n), std::vector<double>(u01, u01 + n));
  });
}

long long lgrngn_b200_philox_call(void *proto)
{
  try { return (long long)slab_of(proto).philox_call; } catch (...) { return -1; }
}

int lgrngn_b200_get_layout(void *proto, unsigned int *sid, unsigned int *ijk, long long cap, long long *n_out)
{
  return guarded_c([&] {
    int64_t n = 0;
    lg::b200::chk(lcx_get_layout(slab_of(proto).e, sid, ijk, cap, &n));
    *n_out = n;
  });
}


long lgrngn_b200_abi_layout(char *buf, long size)
{
  const std::string s = lgrngn_abi_probe::describe_all();
  if (buf && size > 0) { std::strncpy(buf, s.c_str(), size_t(size) - 1); buf[size - 1] = 0; }
  return long(s.size());
}

double lgrngn_b200_common(const char *name, double a, double b, double c, double d, double e)
{
  typedef lcx::cst<double> k;
  const std::string n(name);
  // constants (lib.cpp:40-66)
  if (n == "R_d") return k::R_d();
  if (n == "R_v") return k::R_v();
  if (n == "c_pd") return k::c_pd();
  if (n == "c_pv") return k::c_pv();
  if (n == "c_pw") return k::c_pw();
  if (n == "g") return k::g();
  if (n == "p_1000") return k::p_1000();
  if (n == "eps") return k::eps();
  if (n == "rho_stp") return k::rho_stp();
  if (n == "rho_w") return k::rho_w();
  if (n == "T_tri") return k::T_tri();
  if (n == "p_tri") return k::p_tri();
  if (n == "l_tri") return k::l_tri();
  // functions (common.hpp:20-170)
  if (n == "th_dry2std") return a / std::pow(1 + b * k::R_v() / k::R_d(), k::R_d() / k::c_pd());      // theta_dry.hpp:97-108
  if (n == "th_std2dry") return a * std::pow(1 + b * k::R_v() / k::R_d(), k::R_d() / k::c_pd());      // theta_dry.hpp:83-94
  if (n == "exner") return lcx::exner(a);
  if (n == "p_v") return lcx::p_v(a, b);
  if (n == "p_vs") return lcx::p_vs_cc(a);
  if (n == "p_vs_tet") return lcx::p_vs_tet(a);
  if (n == "r_vs") return lcx::r_vs_cc(a, b);
  if (n == "l_v") return lcx::l_v(a);
  if (n == "T") return lcx::T_of_th_dry(a, b);
  if (n == "p") return lcx::p_of_rhod_rv_T(a, b, c);
  if (n == "visc") return lcx::visc(a);
  if (n == "rw3_cr") return lcx::rw3_cr(a, b, c);
  if (n == "S_cr") return lcx::S_cr(a, b, c);
  if (n == "p_hydro")                                                                                   // hydrostatic.hpp:26-39
  {
    const double R_moist = (k::R_d() + c * k::R_v()) / (1 + c);                                        // moist_air::R(r): moist_air.hpp:55-70
    return k::p_1000() * std::pow(std::pow(e / k::p_1000(), k::R_d() / k::c_pd()) - k::R_d() / k::c_pd() * k::g() / b / R_moist * (a - d),
                                  k::c_pd() / k::R_d());
  }
  if (n == "rhod") return (a - lcx::p_v(a, c)) / (std::pow(a / k::p_1000(), k::R_d() / k::c_pd()) * k::R_d() * b);   // theta_std.hpp:25-32
  return std::numeric_limits<double>::quiet_NaN();
}

}
