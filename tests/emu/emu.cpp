// emu.cpp -- TEST-ONLY host stepping of the kernel phases (etol_b200/csrc/ecuda_phases.cuh).
//
// The container that builds this repo has no GPU. The kernel logic is written as barrier-free
// phases over (tid, nthr); this harness runs those same phase functions on the CPU, looping tid
// where the GPU runs threads and treating each __syncthreads() as the end of a loop. It exists so
// that index/offset logic can be compared with the oracle in the CPU test suite. It is NOT a
// fallback: libecuda.so does not contain it, nothing in etol_b200/ or src/ references it, and the
// GPU parity tests (-m gpu) go through the real kernels via the C ABI.
#include <cstring>
#include <string>
#include <vector>

#include "../../etol_b200/csrc/ecuda_phases.cuh"

using namespace ecuda;

template <int M, int NB>
static void run_c(const ProbDev& pb, const PhaseDev& ph, int p, const EvalIO& io, CtaMem& m, int b, int nthr) {
    for (int t = 0; t < nthr; ++t) phase_c<M, NB>(pb, ph, p, io, m, b, t, nthr);
}

template <int M>
static void run(const ProbDev& pb, const EvalIO& io, int nthr, bool generic) {
    for (int b = 0; b < io.batch; ++b)
        for (int p = 0; p < pb.nphases; ++p) {
            const PhaseDev& ph = pb.ph[p];
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
                for (int t = 0; t < nthr; ++t) phase_b<M>(pb, ph, p, io, m, b, t, nthr);
                switch (generic ? 0 : ph.nb) {  // same dispatch as launch_eval_t
                    case 3: run_c<M, 3>(pb, ph, p, io, m, b, nthr); break;
                    case 4: run_c<M, 4>(pb, ph, p, io, m, b, nthr); break;
                    case 5: run_c<M, 5>(pb, ph, p, io, m, b, nthr); break;
                    default: run_c<M, 0>(pb, ph, p, io, m, b, nthr); break;
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
    std::memset(&pd, 0, sizeof(pd));
    pd.model = desc->model; pd.ns = hp.ns; pd.nc = hp.nc; pd.ne = hp.ne; pd.nphases = hp.nphases;
    pd.nvars = hp.dims.nvars; pd.ncons = hp.dims.ncons; pd.nnz = hp.dims.nnz; pd.nlink = hp.dims.nlinkages;
    pd.linkoff = hp.linkoff; pd.ntracks = desc->ntracks; pd.nway = desc->nwaypoints; pd.track_off = hp.track_off;
    pd.track_size = hp.dims.track_size; pd.rec_size = hp.dims.rec_size; pd.inst_stride = hp.dims.inst_stride;
    pd.maximize = desc->maximize ? 1 : 0; pd.dense = desc->pattern_mode == ECUDA_PATTERN_DENSE_NODE; pd.sf = sf;
    pd.colptr = hp.colptr.data(); pd.isz = isz.data(); pd.sg = sgv.data();
    std::memcpy(pd.xrank, hp.xrank, sizeof(pd.xrank));
    std::memcpy(pd.urank, hp.urank, sizeof(pd.urank));
    std::memcpy(pd.xcnt, hp.xcnt, sizeof(pd.xcnt));
    std::memcpy(pd.ucnt, hp.ucnt, sizeof(pd.ucnt));
    for (int p = 0; p < hp.nphases; ++p) {
        PhaseDev& ph = pd.ph[p];
        ph.N = hp.N[p]; ph.npath = hp.npath[p]; ph.nstat = hp.nstat[p];
        ph.nb = (hp.N[p] + ECUDA_DOT_BLOCK - 1) / ECUDA_DOT_BLOCK;
        ph.zoff = hp.zoff[p]; ph.goff = hp.goff[p]; ph.nvars = hp.nvars_p[p]; ph.inst_off = hp.inst_off[p];
        ph.D = hp.col[p].D.data(); ph.Dt = Dt[p].data(); ph.tau = hp.col[p].tau.data(); ph.w = hp.col[p].w.data();
    }
    std::vector<double> fpart(static_cast<size_t>(desc->batch) * hp.nphases, 0.0);
    EvalIO io{};
    io.x = x; io.inst = inst; io.f = f; io.fpart = fpart.data(); io.g = g; io.jac = jac; io.grad = grad;
    io.jac_mode = jac_mode; io.batch = desc->batch;
    switch (desc->model) {
        case ECUDA_MODEL_SI2D: run<ECUDA_MODEL_SI2D>(pd, io, nthr, generic != 0); break;
        case ECUDA_MODEL_PM3D: run<ECUDA_MODEL_PM3D>(pd, io, nthr, generic != 0); break;
        case ECUDA_MODEL_FW6: run<ECUDA_MODEL_FW6>(pd, io, nthr, generic != 0); break;
        default: return -1;
    }
    return 0;
}
