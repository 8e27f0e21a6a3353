// ROS-free C++ host for libesm_b200.so: the B200-native counterpart of the reference's TensorRT publisher node
// (kitti_publisher/src/kitti_publisher_cuda_node.cpp).  Same stages, same order:
//   preprocess_image (:136-175: pad bottom/right to the network size with black, /255, ImageNet normalise, HWC -> CHW)
//     -> here esm_preprocess_u8_f32 on the device;
//   loadEngine / initializeTensorRT (:177-263: deserialise the .plan, allocate the device buffers)
//     -> here Engine::load: an ".esmeng" file written by esmstereo_b200/engine.py (one recorded forward = a flat list of
//        C-ABI calls + memory plan + weights);
//   H2D (:364-365) -> enqueueV3 (:372) -> sync (:376) -> D2H (:383)
//     -> here cudaMemcpyAsync of the uint8 images, the replayed call list (captured once into a CUDA graph), D2H;
//   crop, cv::medianBlur(5), validity mask, convertTo(CV_16UC1, 256) (:385-404)
//     -> here esm_disparity_publish_u16 on the device.
// No Python, no torch, no allocation after load.  Build: make -C host.   Usage:
//   esm_host <engine.esmeng> <left.u8> <right.u8> <h> <w> <out_prefix> [reps]
// (raw RGB uint8 HWC images of h x w; writes <out_prefix>.disp.f32 (network-size float disparity) and <out_prefix>.u16)
#include <cuda_runtime.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "../include/esm_b200.h"

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      fprintf(stderr, "%s:%d: %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_));       \
      exit(2);                                                                                  \
    }                                                                                           \
  } while (0)

struct Arg {
  char tag;
  int32_t i;
  int64_t q;
  float f;
  void* p;                    // resolved device pointer, or the struct / host blob below
  std::vector<uint8_t> blob;  // 's' / 'h'
};
struct Call {
  int fn;
  std::vector<Arg> a;
};
struct Ref {
  uint32_t seg;
  uint64_t off, n;
};

struct Reader {
  std::vector<uint8_t> d;
  size_t at = 0;
  template <class T>
  T get() {
    T v;
    if (at + sizeof(T) > d.size()) {
      fprintf(stderr, "engine file truncated\n");
      exit(2);
    }
    memcpy(&v, d.data() + at, sizeof(T));
    at += sizeof(T);
    return v;
  }
  const uint8_t* bytes(size_t n) {
    if (at + n > d.size()) {
      fprintf(stderr, "engine file truncated\n");
      exit(2);
    }
    const uint8_t* p = d.data() + at;
    at += n;
    return p;
  }
};

struct Engine {
  std::vector<void*> seg;
  std::vector<std::string> names;
  std::vector<Call> calls;
  float *left = nullptr, *right = nullptr;
  struct Out {
    float* p;
    std::vector<int> shape;
  };
  std::vector<Out> outs;
  std::string meta;
  size_t reserved = 0, state = 0;

  void* at(uint32_t s, uint64_t off) const { return s == 0xFFFFFFFFu ? nullptr : (void*)((uint8_t*)seg[s] + off); }

  void load(const char* path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) {
      fprintf(stderr, "cannot open %s\n", path);
      exit(2);
    }
    Reader r;
    r.d.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
    if (r.d.size() < 8 || memcmp(r.bytes(8), "ESMENG01", 8) != 0) {
      fprintf(stderr, "%s is not an ESMENG01 file\n", path);
      exit(2);
    }
    const uint32_t nseg = r.get<uint32_t>();
    for (uint32_t i = 0; i < nseg; ++i) {
      const uint64_t size = r.get<uint64_t>();
      void* p = nullptr;
      CK(cudaMalloc(&p, size));
      seg.push_back(p);
      reserved += size;
    }
    const uint32_t nblk = r.get<uint32_t>();
    for (uint32_t i = 0; i < nblk; ++i) {
      const uint32_t s = r.get<uint32_t>();
      const uint64_t off = r.get<uint64_t>(), n = r.get<uint64_t>();
      CK(cudaMemcpy(at(s, off), r.bytes(n), n, cudaMemcpyHostToDevice));
      state += n;
    }
    const uint32_t nfn = r.get<uint32_t>();
    for (uint32_t i = 0; i < nfn; ++i) {
      const uint16_t len = r.get<uint16_t>();
      names.emplace_back((const char*)r.bytes(len), len);
    }
    const uint32_t ncall = r.get<uint32_t>();
    calls.resize(ncall);
    for (Call& c : calls) {
      c.fn = r.get<uint16_t>();
      c.a.resize(r.get<uint16_t>());
      for (Arg& a : c.a) {
        a.tag = (char)r.get<uint8_t>();
        a.i = 0; a.q = 0; a.f = 0.f; a.p = nullptr;
        switch (a.tag) {
          case 'i': a.i = r.get<int32_t>(); break;
          case 'q': a.q = r.get<int64_t>(); break;
          case 'f': a.f = r.get<float>(); break;
          case 'S': break;
          case 'p': {
            const uint32_t s = r.get<uint32_t>();
            const uint64_t off = r.get<uint64_t>();
            a.p = at(s, off);
            break;
          }
          case 's': {
            const uint32_t n = r.get<uint32_t>();
            const uint8_t* b = r.bytes(n);
            a.blob.assign(b, b + n);
            const uint16_t nfix = r.get<uint16_t>();
            for (uint16_t k = 0; k < nfix; ++k) {
              const uint32_t where = r.get<uint32_t>(), s = r.get<uint32_t>();
              const uint64_t off = r.get<uint64_t>();
              void* p = at(s, off);
              memcpy(a.blob.data() + where, &p, sizeof(p));
            }
            break;
          }
          case 'h': {
            const uint32_t n = r.get<uint32_t>();
            const uint8_t* b = r.bytes(n);
            a.blob.assign(b, b + n);
            break;
          }
          default: fprintf(stderr, "engine: unknown argument tag %d\n", a.tag); exit(2);
        }
      }
    }
    for (Call& c : calls)
      for (Arg& a : c.a)
        if (a.tag == 's' || a.tag == 'h') a.p = a.blob.data();  // after the vectors have stopped moving
    {
      const uint32_t s = r.get<uint32_t>();
      const uint64_t off = r.get<uint64_t>();
      r.get<uint64_t>();
      left = (float*)at(s, off);
    }
    {
      const uint32_t s = r.get<uint32_t>();
      const uint64_t off = r.get<uint64_t>();
      r.get<uint64_t>();
      right = (float*)at(s, off);
    }
    const uint32_t nout = r.get<uint32_t>();
    for (uint32_t i = 0; i < nout; ++i) {
      Out o;
      const uint32_t s = r.get<uint32_t>();
      const uint64_t off = r.get<uint64_t>();
      o.p = (float*)at(s, off);
      const uint32_t nd = r.get<uint32_t>();
      for (uint32_t k = 0; k < nd; ++k) o.shape.push_back((int)r.get<uint32_t>());
      outs.push_back(o);
    }
    const uint32_t mlen = r.get<uint32_t>();
    meta.assign((const char*)r.bytes(mlen), mlen);
    const uint32_t plen = r.get<uint32_t>();
    const std::string plans((const char*)r.bytes(plen), plen);
    const int np = esm_conv_plans_import(plans.c_str());
    fprintf(stderr, "engine %s: %u segments (%.1f MB reserved), %.1f MB of state, %u calls over %u functions, %d pinned plans\n", path, nseg,
            reserved / 1e6, state / 1e6, ncall, nfn, np);
  }

  // One typed call per C-ABI entry point a forward can contain (include/esm_b200.h).
  int dispatch(const Call& c, void* st) const {
    const std::string& n = names[c.fn];
    const std::vector<Arg>& a = c.a;
#define I(k) a[k].i
#define Q(k) a[k].q
#define F(k) a[k].f
#define P(k) a[k].p
#define FP(k) ((float*)a[k].p)
#define CFP(k) ((const float*)a[k].p)
    if (n == "esm_conv_f32") return esm_conv_f32((const esm_conv_t*)P(0), st);
    if (n == "esm_conv_pf_f32") return esm_conv_pf_f32((const esm_conv_pf_t*)P(0), st);
    if (n == "esm_pf_from_nchw_f32") return esm_pf_from_nchw_f32(CFP(0), Q(1), Q(2), Q(3), Q(4), (const esm_pf_t*)P(5), st);
    if (n == "esm_pf_to_nchw_f32") return esm_pf_to_nchw_f32((const esm_pf_t*)P(0), FP(1), Q(2), Q(3), Q(4), Q(5), st);
    if (n == "esm_copy_f32") return esm_copy_f32(FP(0), CFP(1), Q(2), st);
    if (n == "esm_fill_f32") return esm_fill_f32(FP(0), Q(1), F(2), st);
    if (n == "esm_gwc_volume_f32") return esm_gwc_volume_f32(CFP(0), CFP(1), FP(2), I(3), I(4), I(5), I(6), I(7), I(8), st);
    if (n == "esm_norm_corr_volume_f32") return esm_norm_corr_volume_f32(CFP(0), CFP(1), FP(2), FP(3), I(4), I(5), I(6), I(7), I(8), st);
    if (n == "esm_regression_top2_f32") return esm_regression_top2_f32(CFP(0), FP(1), (int*)P(2), I(3), I(4), I(5), I(6), st);
    if (n == "esm_regression_top2_subpixel_f32")
      return esm_regression_top2_subpixel_f32(CFP(0), Q(1), Q(2), Q(3), Q(4), FP(5), (int*)P(6), I(7), I(8), I(9), I(10), st);
    if (n == "esm_pixel_shuffle3d_f32") return esm_pixel_shuffle3d_f32(CFP(0), Q(1), Q(2), Q(3), Q(4), FP(5), I(6), I(7), I(8), I(9), st);
    if (n == "esm_disparity_regression_f32") return esm_disparity_regression_f32(CFP(0), FP(1), I(2), I(3), I(4), I(5), st);
    if (n == "esm_bilinear_add_f32") return esm_bilinear_add_f32(CFP(0), CFP(1), FP(2), I(3), I(4), I(5), I(6), F(7), st);
    if (n == "esm_sm_pointwise_f32") return esm_sm_pointwise_f32(CFP(0), FP(1), I(2), I(3), I(4), I(5), (const esm_mixer_mlp_t*)P(6), CFP(7), st);
    if (n == "esm_sm_spatial_f32")
      return esm_sm_spatial_f32(CFP(0), FP(1), I(2), I(3), I(4), I(5), CFP(6), CFP(7), I(8), (const esm_mixer_mlp_t*)P(9), CFP(10), st);
    if (n == "esm_sm_layer_f32")
      return esm_sm_layer_f32(CFP(0), FP(1), I(2), I(3), I(4), I(5), (const esm_mixer_mlp_t*)P(6), CFP(7), CFP(8), I(9),
                              (const esm_mixer_mlp_t*)P(10), CFP(11), st);
    if (n == "esm_laf_cost_top7_f32") return esm_laf_cost_top7_f32(CFP(0), FP(1), I(2), I(3), I(4), I(5), st);
    if (n == "esm_laf_attention_f32") return esm_laf_attention_f32(CFP(0), CFP(1), CFP(2), CFP(3), CFP(4), CFP(5), FP(6), I(7), I(8), I(9), I(10), st);
    if (n == "esm_laf_sample_embed_f32")
      return esm_laf_sample_embed_f32(CFP(0), CFP(1), CFP(2), CFP(3), CFP(4), CFP(5), CFP(6), FP(7), I(8), I(9), I(10), I(11), st);
    if (n == "esm_conf_convex_up4_f32") return esm_conf_convex_up4_f32(CFP(0), CFP(1), CFP(2), CFP(3), FP(4), I(5), I(6), I(7), I(8), st);
    if (n == "esm_dwconv2d_f32") return esm_dwconv2d_f32(CFP(0), CFP(1), CFP(2), CFP(3), I(4), FP(5), I(6), I(7), I(8), I(9), I(10), I(11), st);
    if (n == "esm_global_avgpool_f32") return esm_global_avgpool_f32(CFP(0), FP(1), I(2), I(3), I(4), st);
    if (n == "esm_scale_channels_f32") return esm_scale_channels_f32(FP(0), CFP(1), I(2), I(3), I(4), st);
    if (n == "esm_fold_bn_f32") return esm_fold_bn_f32(CFP(0), CFP(1), CFP(2), CFP(3), CFP(4), F(5), I(6), FP(7), FP(8), st);
#undef I
#undef Q
#undef F
#undef P
#undef FP
#undef CFP
    fprintf(stderr, "engine: no dispatcher for %s\n", n.c_str());
    return ESM_ERR_ARG;
  }

  void enqueue(cudaStream_t st) const {
    for (const Call& c : calls) {
      const int rc = dispatch(c, (void*)st);
      if (rc != ESM_OK) {
        fprintf(stderr, "%s failed (rc=%d): %s\n", names[c.fn].c_str(), rc, esm_last_error());
        exit(3);
      }
    }
  }
};

static std::vector<uint8_t> read_file(const char* path, size_t want) {
  std::ifstream f(path, std::ios::binary);
  std::vector<uint8_t> d((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  if (d.size() != want) {
    fprintf(stderr, "%s: expected %zu bytes, got %zu\n", path, want, d.size());
    exit(2);
  }
  return d;
}

int main(int argc, char** argv) {
  if (argc < 7) {
    fprintf(stderr, "usage: %s <engine.esmeng> <left.u8> <right.u8> <h> <w> <out_prefix> [reps]\n", argv[0]);
    return 1;
  }
  const int h = atoi(argv[4]), w = atoi(argv[5]);
  const int reps = argc > 7 ? atoi(argv[7]) : 20;
  Engine eng;
  eng.load(argv[1]);
  if (eng.outs.empty() || eng.outs[0].shape.size() != 3) {
    fprintf(stderr, "engine output 0 is not [B, H, W]\n");
    return 2;
  }
  const int Hp = eng.outs[0].shape[1], Wp = eng.outs[0].shape[2];
  if (h > Hp || w > Wp) {
    fprintf(stderr, "image %d x %d does not fit the engine's %d x %d input\n", h, w, Hp, Wp);
    return 2;
  }
  const std::vector<uint8_t> limg = read_file(argv[2], (size_t)h * w * 3), rimg = read_file(argv[3], (size_t)h * w * 3);
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  uint8_t *d_l, *d_r, *h_l, *h_r;
  uint16_t *d_u16, *h_u16;
  float* h_disp;
  CK(cudaMalloc(&d_l, limg.size()));
  CK(cudaMalloc(&d_r, rimg.size()));
  CK(cudaMalloc(&d_u16, (size_t)h * w * 2));
  CK(cudaMallocHost(&h_l, limg.size()));
  CK(cudaMallocHost(&h_r, rimg.size()));
  CK(cudaMallocHost(&h_u16, (size_t)h * w * 2));
  CK(cudaMallocHost(&h_disp, (size_t)Hp * Wp * 4));
  memcpy(h_l, limg.data(), limg.size());
  memcpy(h_r, rimg.data(), rimg.size());
  const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};

  // the forward, captured once: a replay is one graph launch, like one enqueueV3
  eng.enqueue(st);  // eager once (lazy function attributes, plan lookups)
  CK(cudaStreamSynchronize(st));
  cudaGraph_t graph;
  cudaGraphExec_t exec;
  CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  eng.enqueue(st);
  CK(cudaStreamEndCapture(st, &graph));
  CK(cudaGraphInstantiate(&exec, graph, 0));

  auto frame = [&]() {
    CK(cudaMemcpyAsync(d_l, h_l, limg.size(), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_r, h_r, rimg.size(), cudaMemcpyHostToDevice, st));
    // pad bottom / right with black pixels, normalised like any other (copyMakeBorder before the normalisation, :147-158)
    if (esm_preprocess_u8_f32(d_l, eng.left, 1, h, w, Hp, Wp, 0, 0, 1, mean, stdv, st) != ESM_OK ||
        esm_preprocess_u8_f32(d_r, eng.right, 1, h, w, Hp, Wp, 0, 0, 1, mean, stdv, st) != ESM_OK) {
      fprintf(stderr, "preprocess failed: %s\n", esm_last_error());
      exit(3);
    }
    CK(cudaGraphLaunch(exec, st));
    if (esm_disparity_publish_u16(eng.outs[0].p, d_u16, Hp, Wp, h, w, 192.0f, 256.0f, st) != ESM_OK) {
      fprintf(stderr, "publish failed: %s\n", esm_last_error());
      exit(3);
    }
    CK(cudaMemcpyAsync(h_u16, d_u16, (size_t)h * w * 2, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  };
  for (int i = 0; i < 3; ++i) frame();
  const auto t0 = std::chrono::high_resolution_clock::now();
  for (int i = 0; i < reps; ++i) frame();
  const double ms = std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t0).count() / reps;
  CK(cudaMemcpyAsync(h_disp, eng.outs[0].p, (size_t)Hp * Wp * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  const std::string prefix = argv[6];
  std::ofstream((prefix + ".disp.f32").c_str(), std::ios::binary).write((const char*)h_disp, (size_t)Hp * Wp * 4);
  std::ofstream((prefix + ".u16").c_str(), std::ios::binary).write((const char*)h_u16, (size_t)h * w * 2);
  printf("{\"host\": \"esm_host (C++, no torch)\", \"image\": [%d, %d], \"network\": [%d, %d], \"ms_per_frame\": %.4f, \"frames_per_s\": %.2f, "
         "\"calls_per_frame\": %zu, \"reps\": %d}\n",
         h, w, Hp, Wp, ms, 1e3 / ms, eng.calls.size(), reps);
  return 0;
}
