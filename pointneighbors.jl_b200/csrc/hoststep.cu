// hoststep.cu -- the WCSPH step (update! + interact!) from HOST buffers, pipelined.
//
// What a host-side caller of the reference does per step with a GPU backend is: copy the moved
// coordinates and the state to the device (Adapt / copyto!, benchmarks/run_benchmarks.jl:97-99),
// update!(nhs, ...), interact!, copy dv back.  This object runs that sequence on three streams
// with double-buffered device arrays, so that the host->device copy of step s + 1 and the
// device->host copy of step s - 1 overlap the kernels of step s:
//
//   stream_in  : H2D of (coordinates, v, pressure[, mass]) of step s      -> event in[b]
//   stream_cmp : wait in[b], out[b]; update! (stream-ordered one-pass build) + interact!
//                (gather + tile sweep), nothing synchronised                -> event cmp[b]
//   stream_out : wait cmp[b]; D2H of dv                                      -> event out[b]
//
// pnb_hoststep_submit(s) first enqueues the H2D of step s, then SETTLES step s - 1 (synchronises
// its kernels and reads the error word: a domain error is returned here; a bucket overflow of the
// stream-ordered build makes the library rebuild the cell list, repeat the sweep and the D2H),
// then enqueues the kernels and the D2H of step s.  Host buffers should be pinned
// (pnb_malloc_host) -- pageable memory works but serialises the copies.
#include <chrono>
#include <cstring>

#include "grid.cuh"

struct pnb_hoststep {
    pnb_grid *g;
    int64_t n;
    int nd;
    cudaStream_t s_in, s_cmp, s_out;
    cudaEvent_t ev_in[2], ev_cmp[2], ev_out[2];
    float *y[2], *v[2], *p[2], *dv[2], *mass;
    bool have_mass;
    int64_t steps;           // submitted so far
    bool pending[2];         // kernels of buffer b enqueued, not settled yet
    pnb_wcsph_params prm[2];
    float *dv_host[2];
    double host_s[5];        // host seconds spent in submit: H2D enqueue / settle / update! enqueue / interact! enqueue / D2H enqueue
    bool have_eos;           // pressure computed on the device (compute_pressure!) instead of copied
    float eos_c, eos_rho0, eos_exp, eos_bg;
};

using namespace pnb;

#define HS_CUDA(expr)                                                    \
    do {                                                                 \
        cudaError_t e__ = (expr);                                        \
        if (e__ != cudaSuccess) return ::pnb::cuda_fail(e__, #expr);     \
    } while (0)

extern "C" void pnb_hoststep_destroy(pnb_hoststep *h)
{
    if (!h) return;
    if (h->s_cmp) cudaStreamSynchronize(h->s_cmp);
    if (h->s_out) cudaStreamSynchronize(h->s_out);
    if (h->s_in) cudaStreamSynchronize(h->s_in);
    for (int b = 0; b < 2; b++) {
        cudaFree(h->y[b]); cudaFree(h->v[b]); cudaFree(h->p[b]); cudaFree(h->dv[b]);
        if (h->ev_in[b]) cudaEventDestroy(h->ev_in[b]);
        if (h->ev_cmp[b]) cudaEventDestroy(h->ev_cmp[b]);
        if (h->ev_out[b]) cudaEventDestroy(h->ev_out[b]);
    }
    cudaFree(h->mass);
    if (h->s_in) cudaStreamDestroy(h->s_in);
    if (h->s_cmp) cudaStreamDestroy(h->s_cmp);
    if (h->s_out) cudaStreamDestroy(h->s_out);
    cudaGetLastError();
    delete h;
}

extern "C" pnb_status pnb_hoststep_create(pnb_grid *g, int64_t n, pnb_hoststep **out)
{
    if (!out) { set_error("out is NULL"); return PNB_ERR_ARG; }
    *out = nullptr;
    if (!g || g->f64 || g->hashed) {
        set_error("pnb_hoststep needs a Float32 FullGridCellList search");
        return PNB_ERR_ARG;
    }
    if (n <= 0) { set_error("n must be positive"); return PNB_ERR_ARG; }
    pnb_hoststep *h = new pnb_hoststep();
    memset(h, 0, sizeof(*h));
    h->g = g;
    h->n = n;
    h->nd = g->p.ndims;
    auto fail = [&](cudaError_t e, const char *what) { pnb_status st = cuda_fail(e, what); pnb_hoststep_destroy(h); return st; };
    cudaError_t e;
    if ((e = cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "stream");
    if ((e = cudaStreamCreateWithFlags(&h->s_cmp, cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "stream");
    if ((e = cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "stream");
    const size_t nn = (size_t)n;
    for (int b = 0; b < 2; b++) {
        if ((e = cudaEventCreateWithFlags(&h->ev_in[b], cudaEventDisableTiming)) != cudaSuccess) return fail(e, "event");
        if ((e = cudaEventCreateWithFlags(&h->ev_cmp[b], cudaEventDisableTiming)) != cudaSuccess) return fail(e, "event");
        if ((e = cudaEventCreateWithFlags(&h->ev_out[b], cudaEventDisableTiming)) != cudaSuccess) return fail(e, "event");
        if ((e = cudaMalloc(&h->y[b], sizeof(float) * nn * h->nd)) != cudaSuccess) return fail(e, "cudaMalloc");
        if ((e = cudaMalloc(&h->v[b], sizeof(float) * nn * (h->nd + 1))) != cudaSuccess) return fail(e, "cudaMalloc");
        if ((e = cudaMalloc(&h->p[b], sizeof(float) * nn)) != cudaSuccess) return fail(e, "cudaMalloc");
        if ((e = cudaMalloc(&h->dv[b], sizeof(float) * nn * (h->nd + 1))) != cudaSuccess) return fail(e, "cudaMalloc");
    }
    if ((e = cudaMalloc(&h->mass, sizeof(float) * nn)) != cudaSuccess) return fail(e, "cudaMalloc");
    *out = h;
    return PNB_OK;
}

// StateEquationCole of the system (benchmarks/smoothed_particle_hydrodynamics.jl:64-69): with it a
// submit may pass pressure_host = NULL and the pressure is computed on the device from the density
// row of v, like TrixiParticles.compute_pressure! (:99) does before interact! -- in the reference's
// flow the pressure is derived device state, not a host input.
extern "C" pnb_status pnb_hoststep_set_state_equation(pnb_hoststep *h, float sound_speed,
                                                      float reference_density, float exponent,
                                                      float background_pressure)
{
    if (!h) { set_error("handle is NULL"); return PNB_ERR_ARG; }
    if (!(exponent > 0.0f) || !(reference_density > 0.0f)) {
        set_error("state equation: exponent and reference_density must be positive");
        return PNB_ERR_ARG;
    }
    h->have_eos = true;
    h->eos_c = sound_speed; h->eos_rho0 = reference_density; h->eos_exp = exponent; h->eos_bg = background_pressure;
    return PNB_OK;
}

// settle the kernels of buffer b: error word, repeat after a bucket overflow
static pnb_status hoststep_settle(pnb_hoststep *h, int b)
{
    if (!h->pending[b]) return PNB_OK;
    h->pending[b] = false;
    pnb_status st = check_err_word(h->g, h->s_cmp);          // synchronises stream_cmp
    if (st == PNB_RETRY_INTERNAL) {
        // the cell list has been rebuilt from y[b] (blocking): repeat the sweep and the copy
        st = pnb_wcsph_interact_f32(h->g, h->y[b], h->n, h->y[b], h->n, nullptr, 0, 0, h->v[b], h->v[b],
                                    h->mass, h->mass, h->p[b], h->p[b], &h->prm[b], h->dv[b], h->s_cmp);
        if (st != PNB_OK) return st;
        HS_CUDA(cudaEventRecord(h->ev_cmp[b], h->s_cmp));
        HS_CUDA(cudaStreamWaitEvent(h->s_out, h->ev_cmp[b], 0));
        HS_CUDA(cudaMemcpyAsync(h->dv_host[b], h->dv[b], sizeof(float) * (size_t)h->n * (h->nd + 1),
                                cudaMemcpyDeviceToHost, h->s_out));
        HS_CUDA(cudaEventRecord(h->ev_out[b], h->s_out));
    }
    return st;
}

extern "C" pnb_status pnb_hoststep_wcsph_submit(pnb_hoststep *h, const float *y_host,
                                                const float *v_host, const float *mass_host,
                                                const float *pressure_host,
                                                const pnb_wcsph_params *params, float *dv_host)
{
    if (!h || !y_host || !v_host || !params || !dv_host) {
        set_error("pnb_hoststep_wcsph_submit: NULL argument");
        return PNB_ERR_ARG;
    }
    if (!pressure_host && !h->have_eos) {
        set_error("pressure_host is NULL and no state equation was set (pnb_hoststep_set_state_equation)");
        return PNB_ERR_ARG;
    }
    if (!mass_host && !h->have_mass) { set_error("mass_host is NULL and no mass was given before"); return PNB_ERR_ARG; }
    const int b = (int)(h->steps & 1);
    const size_t nn = (size_t)h->n;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t0 = now();
    // ---- inputs of this step (buffer b is free once the kernels of step s - 2 are done) ----
    HS_CUDA(cudaStreamWaitEvent(h->s_in, h->ev_cmp[b], 0));
    HS_CUDA(cudaMemcpyAsync(h->y[b], y_host, sizeof(float) * nn * h->nd, cudaMemcpyHostToDevice, h->s_in));
    HS_CUDA(cudaMemcpyAsync(h->v[b], v_host, sizeof(float) * nn * (h->nd + 1), cudaMemcpyHostToDevice, h->s_in));
    if (pressure_host)
        HS_CUDA(cudaMemcpyAsync(h->p[b], pressure_host, sizeof(float) * nn, cudaMemcpyHostToDevice, h->s_in));
    if (mass_host) {
        // the mass array is shared by both buffers: the kernels of step s - 1 may still read it
        HS_CUDA(cudaStreamWaitEvent(h->s_in, h->ev_cmp[b ^ 1], 0));
        HS_CUDA(cudaMemcpyAsync(h->mass, mass_host, sizeof(float) * nn, cudaMemcpyHostToDevice, h->s_in));
        h->have_mass = true;
    }
    HS_CUDA(cudaEventRecord(h->ev_in[b], h->s_in));
    double t1 = now();
    h->host_s[0] += t1 - t0;
    // ---- settle the previous step (its kernels overlap the copies just enqueued) ------------
    pnb_status st = hoststep_settle(h, b ^ 1);
    if (st != PNB_OK) return st;
    t0 = now();
    h->host_s[1] += t0 - t1;
    // ---- kernels of this step -----------------------------------------------------------------
    HS_CUDA(cudaStreamWaitEvent(h->s_cmp, h->ev_in[b], 0));
    HS_CUDA(cudaStreamWaitEvent(h->s_cmp, h->ev_out[b], 0));    // dv[b] of step s - 2 has left
    if (!pressure_host) {
        st = pnb_wcsph_compute_pressure_f32(h->nd, h->n, h->v[b], h->eos_c, h->eos_rho0, h->eos_exp, h->eos_bg,
                                            h->p[b], h->s_cmp);
        if (st != PNB_OK) return st;
    }
    st = pnb_grid_build_async_f32(h->g, h->y[b], h->n, h->s_cmp);
    if (st != PNB_OK) return st;
    t1 = now();
    h->host_s[2] += t1 - t0;
    t0 = t1;
    st = pnb_wcsph_interact_async_f32(h->g, h->y[b], h->n, h->v[b], h->mass, h->p[b], params, h->dv[b],
                                      h->s_cmp);
    if (st != PNB_OK) return st;
    HS_CUDA(cudaEventRecord(h->ev_cmp[b], h->s_cmp));
    h->pending[b] = true;
    h->prm[b] = *params;
    h->dv_host[b] = dv_host;
    t1 = now();
    h->host_s[3] += t1 - t0;
    // ---- result ------------------------------------------------------------------------------
    HS_CUDA(cudaStreamWaitEvent(h->s_out, h->ev_cmp[b], 0));
    HS_CUDA(cudaMemcpyAsync(dv_host, h->dv[b], sizeof(float) * nn * (h->nd + 1), cudaMemcpyDeviceToHost,
                            h->s_out));
    HS_CUDA(cudaEventRecord(h->ev_out[b], h->s_out));
    h->host_s[4] += now() - t1;
    h->steps++;
    return PNB_OK;
}

// host seconds spent inside pnb_hoststep_wcsph_submit since the creation, by part: enqueueing the
// H2D copies / waiting for the previous step (settle) / enqueueing update! / enqueueing interact! /
// enqueueing the D2H copy
extern "C" pnb_status pnb_hoststep_host_times(const pnb_hoststep *h, double *out5, int64_t *steps)
{
    if (!h || !out5) { set_error("NULL argument"); return PNB_ERR_ARG; }
    for (int k = 0; k < 5; k++) out5[k] = h->host_s[k];
    if (steps) *steps = h->steps;
    return PNB_OK;
}

extern "C" pnb_status pnb_hoststep_wait(pnb_hoststep *h)
{
    if (!h) { set_error("handle is NULL"); return PNB_ERR_ARG; }
    pnb_status st = hoststep_settle(h, 0);
    if (st != PNB_OK) return st;
    st = hoststep_settle(h, 1);
    if (st != PNB_OK) return st;
    HS_CUDA(cudaStreamSynchronize(h->s_cmp));
    HS_CUDA(cudaStreamSynchronize(h->s_out));
    return PNB_OK;
}
