// emu.cpp -- TEST-ONLY host stepping of the kernel phases (etol_b200/csrc/ecuda_phases.cuh).
//
// The container that builds this repo has no GPU. The kernel logic is written as barrier-free
// phases over (tid, nthr); this harness runs those same phase functions on the CPU, looping tid
// where the GPU runs threads and treating each __syncthreads() as the end of a loop. It exists so
// that index/offset logic can be compared with the oracle in the CPU test suite. It is NOT a
// fallback: libecuda.so does not contain it, nothing in etol_b200/ or src/ references it, and the
// GPU parity tests (-m gpu) go through the real kernels via the C ABI.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../etol_b200/csrc/ecuda_stream.cuh"

using namespace ecuda;

template <int M, int NB>
static void run_c(const ProbDev& pb, const PhaseDev& ph, int p, const EvalIO& io, CtaMem& m, int b, int nthr, int slice,
                  int nslices) {
    for (int t = 0; t < nthr; ++t) phase_c<M, NB>(pb, ph, p, io, m, b, t, nthr, slice, nslices);
}

// the specialised kernel (k_eval_fast): per-thread registers that survive the barrier are an array here
template <int M, int NB, bool FD>
static void run_fast(const ProbDev& pb, const PhaseDev& ph, int p, const EvalIO& io, int b, int nthr) {
    std::vector<double> smem(cta_doubles(pb, ph, nthr, FD ? CARVE_FD : 0), 0.0);
    CtaMem m;
    carve(m, smem.data(), pb, ph, nthr, FD ? CARVE_FD : 0);
    std::memcpy(m.inst, io.inst + static_cast<size_t>(b) * pb.inst_stride, sizeof(double) * pb.inst_stride);
    if (!FD && io.jac)
        for (int t = 0; t < nthr; ++t) fast_copy_template(pb, ph, io, b, t, nthr);
    for (int t = 0; t < nthr; ++t) stage_vars(pb, ph, io, m, b, t, nthr, FD && io.jac != nullptr);
    std::vector<RowRegs<NB>> rr(nthr);
    for (int t = 0; t < nthr; ++t) fast_phase_b<M, NB, FD>(pb, ph, p, io, m, b, t, nthr, rr[t]);
    for (int t = 0; t < nthr; ++t) fast_phase_c<M, NB, FD>(pb, ph, p, io, m, b, t, nthr, rr[t]);
}
// the row-owner kernel (k_eval_rows)
template <int M, int NB, bool FD>
static void run_rows(const ProbDev& pb, const PhaseDev& ph, int p, const EvalIO& io, int b, int nthr) {
    std::vector<double> smem(cta_doubles(pb, ph, nthr, FD ? CARVE_FD : CARVE_ISZ), 0.0);
    CtaMem m;
    carve(m, smem.data(), pb, ph, nthr, FD ? CARVE_FD : CARVE_ISZ);
    std::memcpy(m.inst, io.inst + static_cast<size_t>(b) * pb.inst_stride, sizeof(double) * pb.inst_stride);
    if (!FD && io.jac)
        for (int t = 0; t < nthr; ++t) fast_copy_template(pb, ph, io, b, t, nthr);
    for (int t = 0; t < nthr; ++t) stage_vars<!FD>(pb, ph, io, m, b, t, nthr, FD && io.jac != nullptr);
    std::vector<RowState<M, NB>> rs(nthr);
    for (int t = 0; t < nthr; ++t) {  // no barrier after staging: a thread runs to the end on its own
        rows_values<M, NB, FD>(pb, ph, io, m, b, t, rs[t]);
        if (FD) {
            rows_jacobian<M, NB, FD>(pb, ph, io, m, b, t, rs[t]);
            rows_other<M, NB, FD>(pb, ph, p, io, m, b, t, nthr, rs[t], true, true);
        } else {
            rows_other<M, NB, FD>(pb, ph, p, io, m, b, t, nthr, rs[t], true, false);
        }
    }
    if (!FD)
        for (int t = 0; t < nthr; ++t) {
            rows_jacobian<M, NB, FD>(pb, ph, io, m, b, t, rs[t]);
            rows_other<M, NB, FD>(pb, ph, p, io, m, b, t, nthr, rs[t], false, true);
        }
}
static int g_use_rows = 1;
extern "C" void emu_use_rows(int on) { g_use_rows = on; }
static int g_use_rowsn = 0;
extern "C" void emu_use_rowsn(int on) { g_use_rowsn = on; }
static long g_rowsn_runs = 0;
extern "C" long emu_rowsn_runs() { return g_rowsn_runs; }

// the N-specialised row-owner kernel (k_rows_n_fd): one pass of every thread after staging
template <int M, int N, bool TRK>
static void run_rowsn_fd(const ProbDev& pb, const PhaseDev& ph, int p, const EvalIO& io, int b, int nthr) {
    const bool FD = io.jac && io.jac_mode == ECUDA_JAC_FD_INDEXSET;
    std::vector<double> smem(rn_doubles<M>(pb, N, FD) + 2, 0.0);
    RnMem m;
    rn_carve<M>(m, smem.data(), pb, N, FD);
    CtaMem cm{};
    cm.inst = m.inst;
    cm.z = m.z;
    std::memcpy(m.inst, io.inst + static_cast<size_t>(b) * pb.inst_stride, sizeof(double) * pb.inst_stride);
    for (int t = 0; t < nthr; ++t) {
        if (FD) rn_stage<M, N, true>(pb, ph, io, m, b, t, nthr); else rn_stage<M, N, false>(pb, ph, io, m, b, t, nthr);
    }
    // begin | node groups (the kernel stages each group in a shared-memory ring buffer and stores it with one bulk
    // copy behind a CTA barrier; here the slots are those of the global array itself) | end
    std::vector<RnRow<N>> st(nthr);
    std::vector<double> viol(nthr), fval(nthr);
    for (int t = 0; t < nthr; ++t) {
        if (FD) rn_begin<M, N, true, false>(pb, ph, io, m, cm, b, t, st[t], viol[t], fval[t]);
        else rn_begin<M, N, false, false>(pb, ph, io, m, cm, b, t, st[t], viol[t], fval[t]);
    }
    if (io.jac) {
        double* out = io.jac + static_cast<size_t>(b) * pb.nnz;
        for (int g = 0; g < rn_ngroups<N>(); ++g)
            for (int t = 0; t < nthr; ++t) {
                if (FD) RnGroupRt<M, N, true>::run(g, pb, ph, io, m, b, st[t], out);
                else RnGroupRt<M, N, false>::run(g, pb, ph, io, m, b, st[t], out);
            }
    }
    for (int t = 0; t < nthr; ++t) {
        if (FD) rn_end<M, N, true, TRK, false>(pb, ph, p, io, m, cm, b, t, nthr, st[t], viol[t], fval[t]);
        else rn_end<M, N, false, TRK, false>(pb, ph, p, io, m, cm, b, t, nthr, st[t], viol[t], fval[t]);
    }
    ++g_rowsn_runs;
}
// the streaming exact kernel (k_stream_exact): phase 1 of every thread, (barrier), phase 2 of every thread
template <int M, int N, bool TRK>
static void run_stream(const ProbDev& pb, const PhaseDev& ph, int p, const EvalIO& io, int b, int nthr) {
    std::vector<double> smem(st_doubles<M>(pb, ph, N) + 2, 0.0);
    StMem m;
    st_carve<M>(m, smem.data(), pb, ph, N);
    CtaMem cm{};
    cm.inst = m.inst;
    cm.z = m.z;
    std::memcpy(m.inst, io.inst + static_cast<size_t>(b) * pb.inst_stride, sizeof(double) * pb.inst_stride);
    for (int t = 0; t < nthr; ++t) st_stage<M, N>(pb, ph, io, m, b, t, nthr);
    for (int t = 0; t < nthr; ++t) {
        double viol, fval;
        st_phase1<M, N, TRK, false>(pb, ph, p, io, m, cm, b, t, nthr, viol, fval);
    }
    for (int t = 0; t < nthr; ++t) st_phase2<M, N>(pb, ph, io, m, b, t, nthr);
    ++g_rowsn_runs;
}
static int g_rows_exact = 0;  // exact mode on the row-owner kernel instead of the streaming one (ECUDA_ROWS_EXACT)
extern "C" void emu_rows_exact(int on) { g_rows_exact = on; }
template <int M, int N, bool TRK>
static void run_rowsn(const ProbDev& pb, const PhaseDev& ph, int p, const EvalIO& io, int b, int nthr) {
    const bool fd = io.jac && io.jac_mode == ECUDA_JAC_FD_INDEXSET;
    if (fd || g_rows_exact)
        run_rowsn_fd<M, N, TRK>(pb, ph, p, io, b, nthr);
    else
        run_stream<M, N, TRK>(pb, ph, p, io, b, nthr);
}
// same instantiation list as launch_rows_n (ecuda_api.cu); false: no instantiation, the caller falls back
template <int M>
static bool run_rowsn_any(const ProbDev& pb, const PhaseDev& ph, int p, const EvalIO& io, int b, int nthr) {
    if (nthr < pb.ns * ph.N) return false;
    for (int q = 0; q < pb.nphases; ++q)
        if (pb.ph[q].N != ph.N) return false;
#ifndef ECUDA_USER_MODEL_HEADER
    if constexpr (M == ECUDA_MODEL_PM3D) {
        if (pb.ntracks > 0) return false;
        if (ph.N == 40) return run_rowsn<M, 40, false>(pb, ph, p, io, b, nthr), true;
        if (ph.N == 30) return run_rowsn<M, 30, false>(pb, ph, p, io, b, nthr), true;
    }
    if constexpr (M == ECUDA_MODEL_SI2D) {
        if (ph.N == 33) return run_rowsn<M, 33, true>(pb, ph, p, io, b, nthr), true;
        if (ph.N == 17) return run_rowsn<M, 17, true>(pb, ph, p, io, b, nthr), true;
    }
#endif
    return false;
}

template <int M, int NB>
static void run_fast_mode(const ProbDev& pb, const PhaseDev& ph, int p, const EvalIO& io, int b, int nthr) {
    const bool fd = io.jac && io.jac_mode == ECUDA_JAC_FD_INDEXSET;
    if (g_use_rows) {
        if (fd) run_rows<M, NB, true>(pb, ph, p, io, b, nthr); else run_rows<M, NB, false>(pb, ph, p, io, b, nthr);
    } else {
        if (fd) run_fast<M, NB, true>(pb, ph, p, io, b, nthr); else run_fast<M, NB, false>(pb, ph, p, io, b, nthr);
    }
}

// the persistent exact-mode kernel (k_eval_image): the CTA's image of the phase's triplet range is
// loaded from the template once, phase C overwrites the node-local slots, a memcpy stands for the bulk copy
template <int M, int NB>
static void run_image(const ProbDev& pb, int p, const EvalIO& io, int nthr) {
    const PhaseDev& ph = pb.ph[p];
    std::vector<double> smem(cta_doubles(pb, ph, nthr, 0), 0.0);
    CtaMem m;
    carve(m, smem.data(), pb, ph, nthr, 0);
    const int c0 = pb.colptr[ph.zoff], c1 = pb.colptr[ph.zoff + ph.nvars];
    std::vector<double> image(pb.jtmpl + c0, pb.jtmpl + c1);
    double* vimage = image.data() - c0;
    for (int b = 0; b < io.batch; ++b) {
        std::memcpy(m.inst, io.inst + static_cast<size_t>(b) * pb.inst_stride, sizeof(double) * pb.inst_stride);
        for (int t = 0; t < nthr; ++t) stage_vars(pb, ph, io, m, b, t, nthr, false);
        std::vector<RowRegs<NB>> rr(nthr);
        for (int t = 0; t < nthr; ++t) fast_phase_b<M, NB, false>(pb, ph, p, io, m, b, t, nthr, rr[t]);
        for (int t = 0; t < nthr; ++t) fast_phase_c<M, NB, false, true>(pb, ph, p, io, m, b, t, nthr, rr[t], vimage);
        std::memcpy(io.jac + static_cast<size_t>(b) * pb.nnz + c0, image.data(), sizeof(double) * (c1 - c0));
    }
}

static int g_use_image = 0;
extern "C" void emu_use_image(int on) { g_use_image = on; }
static long g_fast_runs = 0;
extern "C" long emu_fast_runs() { return g_fast_runs; }

static bool fast_ok(const ProbDev& pb, int nthr) {  // same rule as ecuda_set_problem
    int nb = pb.ph[0].nb;
    for (int p = 0; p < pb.nphases; ++p)
        if (pb.ph[p].nb != nb || pb.ns * pb.ph[p].N > nthr) return false;
    return nb >= 3 && nb <= 5;
}

template <int M>
static void run(const ProbDev& pb, const EvalIO& io, int nthr, bool generic) {
    // exact mode with a Jacobian on the fast path: the persistent image kernel (one pass per phase)
    const bool image = g_use_image && !generic && io.jac && io.jac_mode == ECUDA_JAC_EXACT && fast_ok(pb, nthr) && (pb.nnz & 1) == 0;
    if (image) {
        for (int p = 0; p < pb.nphases; ++p) {
            switch (pb.ph[p].nb) {
                case 3: run_image<M, 3>(pb, p, io, nthr); break;
                case 4: run_image<M, 4>(pb, p, io, nthr); break;
                default: run_image<M, 5>(pb, p, io, nthr); break;
            }
        }
    }
    for (int b = 0; b < io.batch; ++b)
        for (int p = 0; p < pb.nphases; ++p) {
            const PhaseDev& ph = pb.ph[p];
            if (image && !io.grad) continue;
            if (image) {  // only the gradient is left (k_grad is a separate launch)
                std::vector<double> smem(cta_doubles(pb, ph, nthr), 0.0);
                CtaMem m;
                carve(m, smem.data(), pb, ph, nthr);
                for (int t = 0; t < nthr; ++t) stage_vars(pb, ph, io, m, b, t, nthr, false);
                for (int t = 0; t < nthr; ++t) cost_nodes<M>(pb, ph, m, t, nthr);
                for (int t = 0; t < nthr; ++t) gradient_phase<M>(pb, ph, io, m, b, t, nthr);
                continue;
            }
            if (!generic && (io.f || io.g || io.jac) && fast_ok(pb, nthr)) {
                ++g_fast_runs;
                if (io.grad) {  // k_grad is a separate launch
                    std::vector<double> smem(cta_doubles(pb, ph, nthr), 0.0);
                    CtaMem m;
                    carve(m, smem.data(), pb, ph, nthr);
                    for (int t = 0; t < nthr; ++t) stage_vars(pb, ph, io, m, b, t, nthr, false);
                    for (int t = 0; t < nthr; ++t) cost_nodes<M>(pb, ph, m, t, nthr);
                    for (int t = 0; t < nthr; ++t) gradient_phase<M>(pb, ph, io, m, b, t, nthr);
                }
                if (g_use_rowsn && run_rowsn_any<M>(pb, ph, p, io, b, nthr)) continue;
                switch (ph.nb) {
                    case 3: run_fast_mode<M, 3>(pb, ph, p, io, b, nthr); break;
                    case 4: run_fast_mode<M, 4>(pb, ph, p, io, b, nthr); break;
                    default: run_fast_mode<M, 5>(pb, ph, p, io, b, nthr); break;
                }
                continue;
            }
            std::vector<double> smem(cta_doubles(pb, ph, nthr), 0.0);
            CtaMem m;
            carve(m, smem.data(), pb, ph, nthr);
            std::memcpy(m.inst, io.inst + static_cast<size_t>(b) * pb.inst_stride, sizeof(double) * pb.inst_stride);
            const bool fd = io.jac && io.jac_mode == ECUDA_JAC_FD_INDEXSET;
            for (int t = 0; t < nthr; ++t) stage_vars(pb, ph, io, m, b, t, nthr, fd);
            if (io.grad) {
                for (int t = 0; t < nthr; ++t) cost_nodes<M>(pb, ph, m, t, nthr);
                for (int t = 0; t < nthr; ++t) gradient_phase<M>(pb, ph, io, m, b, t, nthr);
            }
            if (io.f || io.g || io.jac) {
                // gridDim.y slices of the phase, each its own CTA (fresh shared memory, same staging)
                int nslices = 1;
                for (int q = 0; q < pb.nphases; ++q) nslices = std::max(nslices, generic_slices(pb.ns, pb.ph[q].N, nthr));
                for (int slice = 0; slice < nslices; ++slice) {
                    if (slice > 0) {
                        std::fill(smem.begin(), smem.end(), 0.0);
                        carve(m, smem.data(), pb, ph, nthr);
                        std::memcpy(m.inst, io.inst + static_cast<size_t>(b) * pb.inst_stride,
                                    sizeof(double) * pb.inst_stride);
                        for (int t = 0; t < nthr; ++t) stage_vars(pb, ph, io, m, b, t, nthr, fd);
                    }
                    for (int t = 0; t < nthr; ++t) phase_b<M>(pb, ph, p, io, m, b, t, nthr, slice, nslices);
                    switch (generic ? 0 : ph.nb) {  // same dispatch as launch_eval_t
                        case 3: run_c<M, 3>(pb, ph, p, io, m, b, nthr, slice, nslices); break;
                        case 4: run_c<M, 4>(pb, ph, p, io, m, b, nthr, slice, nslices); break;
                        case 5: run_c<M, 5>(pb, ph, p, io, m, b, nthr, slice, nslices); break;
                        default: run_c<M, 0>(pb, ph, p, io, m, b, nthr, slice, nslices); break;
                    }
                }
            }
        }
    if (io.f && pb.nphases > 1)
        for (int b = 0; b < io.batch; ++b) {
            double tot = io.fpart[static_cast<size_t>(b) * pb.nphases];
            for (int p = 1; p < pb.nphases; ++p) tot = tot + io.fpart[static_cast<size_t>(b) * pb.nphases + p];
            io.f[b] = pb.sf * tot;
        }
}

extern "C" int emu_eval(const ecuda_problem_desc* desc, const double* sz, const double* sg, double sf,
                        const double* inst, const double* x, double* f, double* g, double* jac, double* grad,
                        int jac_mode, int nthr, int generic) {
    HostProblem hp;
    std::string err;
    if (!build_layout(*desc, &hp, &err)) return -1;
    build_structure(&hp);
    hp.col.resize(hp.nphases);
    std::vector<std::vector<double>> Dt(hp.nphases);
    for (int p = 0; p < hp.nphases; ++p) {
        if (!build_collocation(desc->collocation, hp.N[p], &hp.col[p], &err)) return -1;
        size_t N = hp.N[p];
        Dt[p].resize(N * N);
        for (size_t k = 0; k < N; ++k)
            for (size_t l = 0; l < N; ++l) Dt[p][l * N + k] = hp.col[p].D[k * N + l];
    }
    std::vector<double> isz(hp.dims.nvars, 1.0), sgv(hp.dims.ncons, 1.0);
    if (sz)
        for (int c = 0; c < hp.dims.nvars; ++c) isz[c] = 1.0 / sz[c];
    if (sg)
        for (int r = 0; r < hp.dims.ncons; ++r) sgv[r] = sg[r];
    ProbDev pd;
    fill_probdev(hp, &pd);
    pd.sf = sf;
    std::vector<double> tmpl;
    build_jac_template(hp, isz.data(), sgv.data(), &tmpl);
    pd.colptr = hp.colptr.data(); pd.desc = reinterpret_cast<const unsigned long long*>(hp.tdesc.data()); pd.isz = isz.data(); pd.sg = sgv.data(); pd.jtmpl = tmpl.data();
    for (int p = 0; p < hp.nphases; ++p) {
        PhaseDev& ph = pd.ph[p];
        ph.D = hp.col[p].D.data(); ph.Dt = Dt[p].data(); ph.tau = hp.col[p].tau.data(); ph.w = hp.col[p].w.data();
    }
    std::vector<double> fpart(static_cast<size_t>(desc->batch) * hp.nphases, 0.0);
    EvalIO io{};
    io.x = x; io.inst = inst; io.f = f; io.fpart = fpart.data(); io.g = g; io.jac = jac; io.grad = grad;
    io.jac_mode = jac_mode; io.batch = desc->batch;
#ifdef ECUDA_USER_MODEL_HEADER
    // a build of the emulator for one generated user model (tests/emu_binding.py): the id comes from this
    // library's own registry
    if (desc->model < ECUDA_MODEL_USER_BASE) return -1;
    run<ECUDA_MODEL_USER>(pd, io, nthr, generic != 0);
#else
    switch (desc->model) {
        case ECUDA_MODEL_SI2D: run<ECUDA_MODEL_SI2D>(pd, io, nthr, generic != 0); break;
        case ECUDA_MODEL_PM3D: run<ECUDA_MODEL_PM3D>(pd, io, nthr, generic != 0); break;
        case ECUDA_MODEL_FW6: run<ECUDA_MODEL_FW6>(pd, io, nthr, generic != 0); break;
        default: return -1;
    }
#endif
    return 0;
}
