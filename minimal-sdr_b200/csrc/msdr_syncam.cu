// msdr_syncam.cu — K6: synchronous-AM demodulator with PLL, `case SYNCAM` of the demodulation switch,
// Minimal-SDR.ino:631-688 (after wdsp), batched over channels.  Input: the FIR-filtered I and Q streams (what
// msdr_op_fir_fast_q15 / the FIR pair deliver), output: corr[0] narrowed to int16.
//
// A per-sample feedback loop (phase error -> loop filter -> phase) around sinf/cosf/atan2f: serial in time per channel, so one
// lane per channel.  float32 like the reference; the transcendental functions are evaluated in double and rounded to float
// (correctly rounded in all but ~1e-8 of the cases), which agrees with a host libm wherever that one is correctly rounded too.
// Parity is therefore a TOLERANCE (tests/test_gpu_syncam.py), not bit-exactness: libm implementations differ in the last ulp
// (SURVEY 8c).  The loop constants are evaluated on the host with the sketch's own C++ expressions (float exp overload included).
#include "msdr_device.cuh"
#include "msdr_internal.h"
#include "../../include/msdr.h"

#include <cmath>
#include <string>
#include <vector>

namespace msdr {
namespace syncam {

constexpr double kPi = 3.1415926535897932384626433832795; // Arduino.h PI
constexpr int kSampleRate = 6000 * 4;                     // Minimal-SDR.ino:84-85

struct Consts { float omega_min, omega_max, g1, g2; };

// .ino:636-642, the static initialisers verbatim (C++: exp(float) is the float overload)
static Consts make_consts()
{
  typedef float float32_t;
  const float32_t omegaN = 400.0;
  const float32_t zeta = 0.45;
  const float32_t omega_min = 2.0 * kPi * -4000.0 / kSampleRate;
  const float32_t omega_max = 2.0 * kPi * 4000.0 / kSampleRate;
  const float32_t g1 = 1.0 - std::exp(-2.0 * omegaN * zeta / kSampleRate);
  const float32_t g2 = -g1 + 2.0 * (1 - std::exp(-omegaN * zeta / kSampleRate) * cosf(omegaN / kSampleRate * sqrtf(1.0 - zeta * zeta)));
  return Consts{omega_min, omega_max, g1, g2};
}

struct Params {
  const int16_t *I, *Q;
  int16_t *out;
  size_t stride, ostride;
  uint32_t C, n;        // rows; samples per row, multiple of 8
  float *state;         // [3][Cpad]: fil_out, omega2, phzerror
  uint32_t Cpad;
  const uint32_t *chmap; // row -> state index (NULL: identity); used by the receive chain for its SYNCAM channels
  const uint8_t *row_sel; // optional: only rows with row_sel[row] == 255 are processed
  Consts k;
};

__global__ void __launch_bounds__(128) syncam_kernel(const Params p)
{
  const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= p.C || (p.row_sel && p.row_sel[row] != 255)) return;
  const uint32_t ch = p.chmap ? p.chmap[row] : row;
  float fil_out = p.state[ch], omega2 = p.state[p.Cpad + ch], phzerror = p.state[2 * (size_t)p.Cpad + ch];
  const uint4 *pi = reinterpret_cast<const uint4 *>(p.I + (size_t)row * p.stride);
  const uint4 *pq = reinterpret_cast<const uint4 *>(p.Q + (size_t)row * p.stride);
  uint4 *po = reinterpret_cast<uint4 *>(p.out + (size_t)row * p.ostride);
  const double two_pi = 2.0 * kPi;
  for (uint32_t v = 0; v < p.n / 8; ++v) {
    const uint4 vi = pi[v], vq = pq[v];
    const uint32_t wi[4] = {vi.x, vi.y, vi.z, vi.w}, wq[4] = {vq.x, vq.y, vq.z, vq.w};
    uint32_t wo[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int o[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float fi = (float)(short)((wi[k] >> (16 * h)) & 0xFFFFu), fq = (float)(short)((wq[k] >> (16 * h)) & 0xFFFFu);
        double sd, cd;
        sincos((double)phzerror, &sd, &cd);
        const float Sin = __double2float_rn(sd), Cos = __double2float_rn(cd);
        const float ai = __fmul_rn(Cos, fi), bi = __fmul_rn(Sin, fi), aq = __fmul_rn(Cos, fq), bq = __fmul_rn(Sin, fq);
        const float corr0 = __fadd_rn(ai, bq), corr1 = __fadd_rn(-bi, aq);
        o[h] = (int)(short)__float2int_rz(corr0);
        const float det = __double2float_rn(atan2((double)corr1, (double)corr0));
        const float del_out = fil_out;
        omega2 = __fadd_rn(omega2, __fmul_rn(p.k.g2, det));
        if (omega2 < p.k.omega_min) omega2 = p.k.omega_min;
        else if (omega2 > p.k.omega_max) omega2 = p.k.omega_max;
        fil_out = __fadd_rn(__fmul_rn(p.k.g1, det), omega2);
        phzerror = __fadd_rn(phzerror, del_out);
        while ((double)phzerror >= two_pi) phzerror = __double2float_rn(__dsub_rn((double)phzerror, two_pi)); // wrap round 2 PI
        while (phzerror < 0.0f) phzerror = __double2float_rn(__dadd_rn((double)phzerror, two_pi));
      }
      wo[k] = ((uint32_t)o[0] & 0xFFFFu) | ((uint32_t)o[1] << 16);
    }
    po[v] = make_uint4(wo[0], wo[1], wo[2], wo[3]);
  }
  p.state[ch] = fil_out; p.state[p.Cpad + ch] = omega2; p.state[2 * (size_t)p.Cpad + ch] = phzerror;
}

} // namespace syncam

// rows of filtered I/Q -> audio, PLL state indexed through chmap (msdr_capi.cu: SYNCAM channels of a receive chain)
cudaError_t launch_syncam(const int16_t *I, const int16_t *Q, size_t stride, int16_t *out, size_t ostride, uint32_t rows, uint32_t n, float *state,
                          uint32_t Cpad, const uint32_t *chmap, const uint8_t *row_sel, cudaStream_t s)
{
  if (rows == 0 || n == 0) return cudaSuccess;
  static const syncam::Consts k = syncam::make_consts();
  syncam::Params p{};
  p.I = I; p.Q = Q; p.out = out; p.stride = stride; p.ostride = ostride; p.C = rows; p.n = n; p.state = state; p.Cpad = Cpad; p.chmap = chmap; p.row_sel = row_sel; p.k = k;
  syncam::syncam_kernel<<<(rows + 127) / 128, 128, 0, s>>>(p);
  return cudaGetLastError();
}
} // namespace msdr

struct msdr_syncam {
  int device = 0;
  uint32_t C = 0, Cpad = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  float *d_state = nullptr;
  int16_t *d_buf = nullptr; // staging for host updates: I, Q, out
  size_t stage_samples = 0;
  msdr::syncam::Consts k{};
  uint64_t launches = 0;
  std::string err;
};

namespace {
thread_local std::string g_sc_error;
int sc_fail(msdr_syncam *s, int code, const std::string &msg) { if (s) s->err = msg; else g_sc_error = msg; return code; }
#define SCK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return sc_fail(sc, MSDR_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); } while (0)
}

extern "C" {

const char *msdr_syncam_last_error(const msdr_syncam *sc) { return sc ? sc->err.c_str() : g_sc_error.c_str(); }

int msdr_syncam_create(msdr_syncam **out, int device, uint32_t n_channels)
{
  if (!out || n_channels == 0) return sc_fail(nullptr, MSDR_ERR_ARGUMENT, "syncam_create: bad arguments");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return sc_fail(nullptr, MSDR_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
  if (device < 0 || device >= ndev) return sc_fail(nullptr, MSDR_ERR_ARGUMENT, "syncam_create: bad device");
  msdr_syncam *sc = new msdr_syncam();
  sc->device = device; sc->C = n_channels; sc->Cpad = (n_channels + 31u) & ~31u;
  sc->k = msdr::syncam::make_consts();
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&sc->own_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaMalloc(&sc->d_state, (size_t)3 * sc->Cpad * 4);
  if (e == cudaSuccess) e = cudaMemset(sc->d_state, 0, (size_t)3 * sc->Cpad * 4); // fil_out = omega2 = phzerror = 0, .ino:643-645
  if (e == cudaSuccess) e = cudaDeviceSynchronize(); // legacy-stream memset vs this object's non-blocking stream
  if (e != cudaSuccess) {
    g_sc_error = std::string("syncam_create: ") + cudaGetErrorString(e);
    cudaFree(sc->d_state);
    if (sc->own_stream) cudaStreamDestroy(sc->own_stream);
    delete sc;
    return MSDR_ERR_CUDA;
  }
  sc->stream = sc->own_stream;
  *out = sc;
  return MSDR_OK;
}

void msdr_syncam_destroy(msdr_syncam *sc)
{
  if (!sc) return;
  cudaSetDevice(sc->device);
  if (sc->stream) cudaStreamSynchronize(sc->stream);
  cudaFree(sc->d_state); cudaFree(sc->d_buf);
  if (sc->own_stream) cudaStreamDestroy(sc->own_stream);
  delete sc;
}

int msdr_syncam_set_stream(msdr_syncam *sc, void *cuda_stream)
{
  if (!sc) return MSDR_ERR_ARGUMENT;
  sc->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : sc->own_stream;
  return MSDR_OK;
}

int msdr_syncam_update_device(msdr_syncam *sc, const int16_t *d_I, const int16_t *d_Q, int16_t *d_out, uint32_t n_blocks, size_t stride)
{
  if (!sc) return MSDR_ERR_ARGUMENT;
  if (n_blocks == 0) return MSDR_OK;
  if (!d_I || !d_Q || !d_out || (uint64_t)n_blocks * MSDR_BLOCK_SAMPLES > stride) return sc_fail(sc, MSDR_ERR_ARGUMENT, "syncam_update: bad buffers / stride < n_blocks*128");
  if (((uintptr_t)d_I & 15u) || ((uintptr_t)d_Q & 15u) || ((uintptr_t)d_out & 15u) || (stride & 7u))
    return sc_fail(sc, MSDR_ERR_ARGUMENT, "syncam_update_device: buffers must be 16-byte aligned and stride a multiple of 8 samples");
  SCK(cudaSetDevice(sc->device));
  msdr::syncam::Params p{};
  p.I = d_I; p.Q = d_Q; p.out = d_out; p.stride = stride; p.ostride = stride; p.chmap = nullptr; p.row_sel = nullptr; p.C = sc->C; p.n = n_blocks * MSDR_BLOCK_SAMPLES; p.state = sc->d_state; p.Cpad = sc->Cpad; p.k = sc->k;
  msdr::syncam::syncam_kernel<<<(sc->C + 127) / 128, 128, 0, sc->stream>>>(p);
  SCK(cudaGetLastError());
  sc->launches++;
  return MSDR_OK;
}

int msdr_syncam_update(msdr_syncam *sc, const int16_t *I, const int16_t *Q, int16_t *out, uint32_t n_blocks, size_t stride)
{
  if (!sc) return MSDR_ERR_ARGUMENT;
  if (n_blocks == 0) return MSDR_OK;
  if (!I || !Q || !out || (uint64_t)n_blocks * MSDR_BLOCK_SAMPLES > stride) return sc_fail(sc, MSDR_ERR_ARGUMENT, "syncam_update: bad buffers / stride < n_blocks*128");
  SCK(cudaSetDevice(sc->device));
  const size_t L = (size_t)n_blocks * MSDR_BLOCK_SAMPLES, need = (size_t)sc->C * L;
  if (need > sc->stage_samples) {
    SCK(cudaStreamSynchronize(sc->stream));
    cudaFree(sc->d_buf);
    sc->d_buf = nullptr; sc->stage_samples = 0;
    SCK(cudaMalloc(&sc->d_buf, need * 2 * 3));
    sc->stage_samples = need;
  }
  int16_t *dI = sc->d_buf, *dQ = sc->d_buf + need, *dO = sc->d_buf + 2 * need;
  SCK(cudaMemcpy2DAsync(dI, L * 2, I, stride * 2, L * 2, sc->C, cudaMemcpyHostToDevice, sc->stream));
  SCK(cudaMemcpy2DAsync(dQ, L * 2, Q, stride * 2, L * 2, sc->C, cudaMemcpyHostToDevice, sc->stream));
  SCK(cudaDeviceSynchronize()); // the copies ran in the legacy stream; this object's stream is non-blocking
  int st = msdr_syncam_update_device(sc, dI, dQ, dO, n_blocks, L);
  if (st != MSDR_OK) return st;
  SCK(cudaMemcpy2DAsync(out, stride * 2, dO, L * 2, L * 2, sc->C, cudaMemcpyDeviceToHost, sc->stream));
  SCK(cudaStreamSynchronize(sc->stream));
  return MSDR_OK;
}

int msdr_syncam_get_state(msdr_syncam *sc, uint32_t ch, float *fil_out, float *omega2, float *phzerror)
{
  if (!sc || !fil_out || !omega2 || !phzerror || ch >= sc->C) return MSDR_ERR_ARGUMENT;
  SCK(cudaSetDevice(sc->device));
  SCK(cudaStreamSynchronize(sc->stream));
  SCK(cudaMemcpy(fil_out, sc->d_state + ch, 4, cudaMemcpyDeviceToHost));
  SCK(cudaMemcpy(omega2, sc->d_state + sc->Cpad + ch, 4, cudaMemcpyDeviceToHost));
  SCK(cudaMemcpy(phzerror, sc->d_state + 2 * (size_t)sc->Cpad + ch, 4, cudaMemcpyDeviceToHost));
  return MSDR_OK;
}

int msdr_syncam_set_state(msdr_syncam *sc, uint32_t ch, float fil_out, float omega2, float phzerror)
{
  if (!sc || ch >= sc->C) return MSDR_ERR_ARGUMENT;
  SCK(cudaSetDevice(sc->device));
  SCK(cudaStreamSynchronize(sc->stream));
  SCK(cudaMemcpy(sc->d_state + ch, &fil_out, 4, cudaMemcpyHostToDevice));
  SCK(cudaMemcpy(sc->d_state + sc->Cpad + ch, &omega2, 4, cudaMemcpyHostToDevice));
  SCK(cudaMemcpy(sc->d_state + 2 * (size_t)sc->Cpad + ch, &phzerror, 4, cudaMemcpyHostToDevice));
  SCK(cudaDeviceSynchronize()); // the copies ran in the legacy stream; this object's stream is non-blocking
  return MSDR_OK;
}

/* the loop constants as the object uses them (for tests: they must equal the reference's static initialisers) */
int msdr_syncam_constants(float *omega_min, float *omega_max, float *g1, float *g2)
{
  if (!omega_min || !omega_max || !g1 || !g2) return MSDR_ERR_ARGUMENT;
  const msdr::syncam::Consts k = msdr::syncam::make_consts();
  *omega_min = k.omega_min; *omega_max = k.omega_max; *g1 = k.g1; *g2 = k.g2;
  return MSDR_OK;
}

uint64_t msdr_syncam_launch_count(const msdr_syncam *sc) { return sc ? sc->launches : 0; }

} // extern "C"
