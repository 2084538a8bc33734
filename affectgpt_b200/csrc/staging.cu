// Input staging on the GPU: the two steps that sit between the decoders and the encoders and that the reference runs
// on the CPU for every sample.
//
//  * face crops: ``load_face`` (A/my_affectgpt/processors/video_processor.py:262-296) = np.load of the OpenFace crop
//    stack [N,h,w,3] u8, ``cv2.resize(face, (224, 224))`` of EVERY crop, then the uniform index rule.  Here only the
//    sampled crops are resized, straight into the planar u8 layout the ViT patchify kernel reads.  The arithmetic is
//    OpenCV's fixed-point INTER_LINEAR for 8-bit images, bit for bit (resize.cpp: 11-bit coefficients from float
//    source coordinates, horizontal pass in 32-bit, vertical pass ((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2 >> 2;
//    horizontal border taps reset the weight to (2048, 0), vertical border taps keep their weights and clamp the row).
//    Verified against cv2 itself through the reference's own load_face: tests/golden/reference_face_resize.npz.
//  * audio: ``load_audio`` (A/my_affectgpt/models/ImageBind/data.py:188-212) after decode/resample: stereo -> mean,
//    zero padding to 2 s, eight [start, start + 32000) slices.  The slice starts are computed on the host in exact
//    rational arithmetic (affectgpt_b200/sampling.py); the kernel is the gather.
#include "agb_internal.h"

namespace agb {

namespace {

struct Tap {
  int i0, i1, a0, a1;
};

// OpenCV resize.cpp, INTER_LINEAR, fixed point: source tap and 11-bit weights of destination coordinate d
__device__ __forceinline__ Tap cv_linear_tap(int d, int src, int dst, bool reset_border_weight) {
  const double inv_scale = (double)dst / (double)src;
  const double scale = 1.0 / inv_scale;
  // (float)((d + 0.5) * scale - 0.5) with separately rounded multiply and subtract (no FMA contraction)
  float f = (float)__dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5);
  int s = (int)floorf(f);
  f = __fsub_rn(f, (float)s);
  if (reset_border_weight) {
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= src - 1) { f = 0.f; s = src - 1; }
  }
  Tap t;
  t.i0 = min(max(s, 0), src - 1);
  t.i1 = min(max(s + 1, 0), src - 1);
  t.a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));    // saturate_cast<short>(cvRound(...)); |v| <= 2048
  t.a1 = __float2int_rn(__fmul_rn(f, 2048.f));
  return t;
}

// grid (224 rows, T frames), block 224: out[t, c, dy, dx] <- faces[idx[t]] resized
// A crop index outside [0, n_src) is a caller error the C ABI cannot see before the launch (the indices live in device
// memory): such a frame is written as zeros and flagged in *bad (checked by agb_stage_faces after the launch).
__global__ void __launch_bounds__(224) face_resize_kernel(const uint8_t* __restrict__ faces, int n_src, int h, int w,
                                                         const int* __restrict__ idx, uint8_t* __restrict__ out,
                                                         int* __restrict__ bad) {
  const int dy = blockIdx.x, t = blockIdx.y, dx = threadIdx.x;
  const int src = idx[t];
  if (src < 0 || src >= n_src) {
    if (dy == 0 && dx == 0) atomicAdd(bad, 1);
#pragma unroll
    for (int c = 0; c < 3; ++c) out[(((long long)t * 3 + c) * 224 + dy) * 224 + dx] = 0;
    return;
  }
  const uint8_t* img = faces + (long long)src * h * w * 3;
  const Tap ty = cv_linear_tap(dy, h, 224, false);
  const Tap tx = cv_linear_tap(dx, w, 224, true);
  const uint8_t* r0 = img + (long long)ty.i0 * w * 3;
  const uint8_t* r1 = img + (long long)ty.i1 * w * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int s0 = r0[tx.i0 * 3 + c] * tx.a0 + r0[tx.i1 * 3 + c] * tx.a1;     // horizontal pass, scaled by 2^11
    const int s1 = r1[tx.i0 * 3 + c] * tx.a0 + r1[tx.i1 * 3 + c] * tx.a1;
    const int v = (((ty.a0 * (s0 >> 4)) >> 16) + ((ty.a1 * (s1 >> 4)) >> 16) + 2) >> 2;
    out[(((long long)t * 3 + c) * 224 + dy) * 224 + dx] = (uint8_t)min(max(v, 0), 255);
  }
}

// out[clip, i] = sample (start[clip] + i) of the mono mix, zero beyond the end of the waveform
__global__ void __launch_bounds__(256) audio_gather_kernel(const float* __restrict__ wav, int ch, long long n,
                                                          const long long* __restrict__ starts, int clip_len,
                                                          float* __restrict__ out) {
  const int clip = blockIdx.y;
  const long long s0 = starts[clip];
  for (int i = blockIdx.x * 256 + threadIdx.x; i < clip_len; i += gridDim.x * 256) {
    const long long p = s0 + i;
    float v = 0.f;
    if (p >= 0 && p < n) v = (ch == 2) ? (wav[p] + wav[n + p]) * 0.5f : wav[p];    // data.py:190-191 (two channels only)
    out[(long long)clip * clip_len + i] = v;
  }
}

// Polyphase sinc resampling (torchaudio.functional.resample, called by load_audio at ImageBind/data.py:186-187):
//   y[c, j * new + p] = sum_k bank[p, k] * xpad[c, j * orig + k],   xpad[i] = x[i - width] (zero outside [0, n)),
// bank [new, taps] with taps = 2 * width + orig (built on the host with torchaudio's formula).  One thread per output
// sample, k ascending in fp32; the phase rows of a block's outputs sit in shared memory when they fit.
__global__ void __launch_bounds__(256) resample_kernel(const float* __restrict__ x, long long n, const float* __restrict__ bank,
                                                       int orig, int nw, int width, int taps, float* __restrict__ y,
                                                       long long n_out) {
  const int c = blockIdx.y;
  const float* xc = x + (long long)c * n;
  for (long long o = (long long)blockIdx.x * 256 + threadIdx.x; o < n_out; o += (long long)gridDim.x * 256) {
    const long long j = o / nw;
    const int p = (int)(o - j * nw);
    const float* bp = bank + (long long)p * taps;
    const long long base = j * orig - width;
    float acc = 0.f;
    for (int k = 0; k < taps; ++k) {
      const long long i = base + k;
      const float v = (i >= 0 && i < n) ? __ldg(xc + i) : 0.f;
      acc = fmaf(__ldg(bp + k), v, acc);
    }
    y[(long long)c * n_out + o] = acc;
  }
}

}  // namespace

int resample_f32(const float* wav, int ch, int64_t n, const float* bank, int orig, int nw, int width, int taps, float* out,
                 int64_t n_out, cudaStream_t s) {
  AGB_REQUIRE(wav && bank && out && ch >= 1 && n > 0 && n_out > 0, "resample: bad arguments");
  AGB_REQUIRE(orig > 0 && nw > 0 && width > 0 && taps == 2 * width + orig, "resample: bad filter bank shape");
  LaunchTimer timer(PROF_FRONTEND, 2.0 * ch * n_out * taps, 4.0 * ch * (n + n_out), s);
  long long blocks = (n_out + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  resample_kernel<<<dim3((unsigned)blocks, (unsigned)ch), 256, 0, s>>>(wav, n, bank, orig, nw, width, taps, out, n_out);
  AGB_CHECK_LAUNCH();
  return 0;
}

int stage_faces_u8(const uint8_t* faces, int64_t n_src, int h, int w, const int* idx_dev, int T, uint8_t* out,
                   cudaStream_t s) {
  AGB_REQUIRE(faces && idx_dev && out && n_src > 0 && n_src < (1ll << 31) && h > 0 && w > 0 && T > 0,
              "stage_faces: bad arguments");
  static int* bad_by_device[64] = {};       // one 4-byte counter per device (staging is not on the graph-captured path)
  int dev = 0;
  AGB_CHECK_CUDA(cudaGetDevice(&dev));
  AGB_REQUIRE(dev >= 0 && dev < 64, "stage_faces: device index out of range");
  if (bad_by_device[dev] == nullptr) AGB_CHECK_CUDA(cudaMalloc(&bad_by_device[dev], sizeof(int)));
  int* bad_dev = bad_by_device[dev];
  AGB_CHECK_CUDA(cudaMemsetAsync(bad_dev, 0, sizeof(int), s));
  {
    LaunchTimer timer(PROF_FRONTEND, 0.0, (double)T * (4.0 * 3 * 224 * 224 + 3.0 * 224 * 224), s);
    face_resize_kernel<<<dim3(224, T), 224, 0, s>>>(faces, (int)n_src, h, w, idx_dev, out, bad_dev);
    AGB_CHECK_LAUNCH();
  }
  int bad = 0;
  AGB_CHECK_CUDA(cudaMemcpyAsync(&bad, bad_dev, sizeof(int), cudaMemcpyDeviceToHost, s));
  AGB_CHECK_CUDA(cudaStreamSynchronize(s));
  AGB_REQUIRE(bad == 0, "stage_faces: a crop index lies outside [0, n_src)");
  return 0;
}

int stage_audio_f32(const float* wav, int ch, int64_t n, const long long* starts_dev, int clips, int clip_len, float* out,
                    cudaStream_t s) {
  AGB_REQUIRE(wav && starts_dev && out && ch >= 1 && n >= 0 && clips > 0 && clip_len > 0, "stage_audio: bad arguments");
  // load_audio averages exactly two channels and keeps any other channel count as it is (data.py:190-191); the
  // [clips, 1, clip_len] output of this entry point can only represent the mono and the stereo case
  AGB_REQUIRE(ch <= 2, "stage_audio: more than two channels are not supported (mix down on the caller side)");
  LaunchTimer timer(PROF_FRONTEND, 0.0, (double)clips * clip_len * 4.0 * (ch == 2 ? 3.0 : 2.0), s);
  audio_gather_kernel<<<dim3(32, clips), 256, 0, s>>>(wav, ch, n, starts_dev, clip_len, out);
  AGB_CHECK_LAUNCH();
  return 0;
}

}  // namespace agb
