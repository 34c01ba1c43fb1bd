// Host side of the layers outside the fused MobileNetV2 path (included by synergy_b200.cu):
//   * PointNet refinement heads MLP_for / MLP_rev (reference backbone_nets/pointnet_backbone.py:7-106) and the
//     training-forward losses (loss_definition.py:8-42) -- SURVEY.md section 8 rows a10 / f4;
// all on tc_gemm_kernel (kernels_gemm.cuh) with BatchNorm folded in float64 at commit time.

namespace {

struct GemmLayer {
  int K = 0, N = 0, Kp = 0, nr = 0, nranges = 0, act = kActNone;
  int ksize = 0, stride = 1, pad = 0;
  uint8_t* d_img = nullptr;
  float *d_bias = nullptr, *d_osc = nullptr;
  float* d_wkn = nullptr;            // small-K layers only: [K][N] fp32 for the CUDA-core kernel
};

struct RawLayer {                    // what the setters keep until commit: conv weight (cout, cin*k*k), conv bias, BN
  std::vector<float> w, b, g, beta, m, v;
  int cout = 0, cin = 0, ksize = 1;
  float eps = 1e-5f;
  bool has_bn = false, set = false;
};

void free_layer(GemmLayer& L) {
  cudaFree(L.d_img); cudaFree(L.d_bias); cudaFree(L.d_osc); cudaFree(L.d_wkn);
  L = GemmLayer{};
}

// BatchNorm (eval) folded into the conv: w'[n][k] = w[n][k] * s[n], b'[n] = (b[n] - mean[n]) * s[n] + beta[n]
void fold_bn(const RawLayer& r, std::vector<double>& w, std::vector<double>& b) {
  const int kk = r.cin * r.ksize * r.ksize;
  w.resize((size_t)r.cout * kk);
  b.resize(r.cout);
  for (int n = 0; n < r.cout; ++n) {
    const double s = r.has_bn ? (double)r.g[n] / sqrt((double)r.v[n] + (double)r.eps) : 1.0;
    for (int k = 0; k < kk; ++k) w[(size_t)n * kk + k] = (double)r.w[(size_t)n * kk + k] * s;
    const double cb = r.b.empty() ? 0.0 : (double)r.b[n];
    b[n] = r.has_bn ? (cb - (double)r.m[n]) * s + (double)r.beta[n] : cb;
  }
}

// Wnk: folded weights [N][K] (K already in the GEMM's k order); builds the device image of one layer.
int build_gemm_layer(GemmLayer& L, const std::vector<double>& Wnk, const std::vector<double>& bias, int N, int K, int act) {
  free_layer(L);
  L.N = N; L.K = K; L.act = act;
  L.Kp = (K + 15) / 16 * 16;
  L.nr = std::min(kGmMaxNr, (N + 15) / 16 * 16);
  L.nranges = (N + L.nr - 1) / L.nr;
  std::vector<uint8_t> img((size_t)L.nranges * L.nr * L.Kp * 4, 0);
  std::vector<float> osc((size_t)L.nranges * L.nr, 0.f), bs((size_t)L.nranges * L.nr, 0.f);
  std::vector<float> rowf(K);
  uint16_t* base = reinterpret_cast<uint16_t*>(img.data());
  const size_t lbo = (size_t)(L.nr / 8) * 128;
  for (int n = 0; n < N; ++n) {
    for (int k = 0; k < K; ++k) rowf[k] = (float)Wnk[(size_t)n * K + k];
    const float ws = channel_scale(rowf.data(), 1, K);
    osc[n] = 1.0f / ws;
    bs[n] = (float)bias[n];
    const int j = n / L.nr, nl = n % L.nr;
    for (int k0 = 0; k0 < L.Kp; k0 += kGmKC) {
      const int kc = std::min(kGmKC, L.Kp - k0);
      uint16_t* hi = base + ((size_t)j * L.nr * L.Kp * 4 + (size_t)L.nr * k0 * 4) / 2;
      uint16_t* lo = hi + (size_t)L.nr * kc;
      for (int kl = 0; kl < kc && k0 + kl < K; ++kl) {
        const size_t off = ((size_t)(nl / 8) * 128 + (size_t)(kl / 8) * lbo + (nl % 8) * 16 + (kl % 8) * 2) / 2;
        split_f16_host(rowf[k0 + kl] * ws, hi[off], lo[off]);
      }
    }
  }
  SYN_CUDA(cudaMalloc(&L.d_img, img.size()));
  SYN_CUDA(cudaMemcpy(L.d_img, img.data(), img.size(), cudaMemcpyHostToDevice));
  int rc = upload(&L.d_osc, osc);
  if (rc != SYN_OK) return rc;
  return upload(&L.d_bias, bs);
}

struct GemmIO {
  const float* A = nullptr; int lda = 0; int M = 0;
  const unsigned* rowmax_in = nullptr;
  float* out = nullptr;
  unsigned* rowmax_out = nullptr;
  const float* addend = nullptr; int addend_group = 1;
  const float* residual = nullptr;
  unsigned* colmax_out = nullptr; int colmax_group = 1;
  int H = 0, W = 0, C = 0, HO = 0, WO = 0;     // conv mode
};

int launch_gemm(syn_handle* h, const GemmLayer& L, const GemmIO& io, cudaStream_t st, const char* name) {
  static bool attr_set[16] = {};
  if (!attr_set[h->device & 15]) {
    SYN_CUDA(cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGmSmem));
    attr_set[h->device & 15] = true;
  }
  if (io.rowmax_out != nullptr) SYN_CUDA(cudaMemsetAsync(io.rowmax_out, 0, (size_t)io.M * sizeof(unsigned), st));
  GemmArgs a;
  a.A = io.A; a.Wimg = L.d_img; a.bias = L.d_bias; a.oscale = L.d_osc; a.addend = io.addend; a.residual = io.residual;
  a.out = io.out; a.rowmax_in = io.rowmax_in; a.rowmax_out = io.rowmax_out; a.colmax_out = io.colmax_out;
  a.addend_group = io.addend_group; a.colmax_group = io.colmax_group;
  a.M = io.M; a.K = L.K; a.N = L.N; a.Kp = L.Kp; a.nr = L.nr; a.lda = io.lda; a.act = L.act;
  a.ksize = L.ksize; a.stride = L.stride; a.pad = L.pad; a.H = io.H; a.W = io.W; a.C = io.C; a.HO = io.HO; a.WO = io.WO;
  a.err = h->d_err;
  dim3 grid((io.M + 127) / 128, L.nranges);
  tc_gemm_kernel<<<grid, kGmThreads, kGmSmem, st>>>(a);
  SYN_LAUNCH_CHECK(name);
  mark(h, st, name);
  return SYN_OK;
}

constexpr int kPts = 68;
constexpr int kFaceVecLd = 2360;     // 1024 + 1280 + 40 + 10 = 2354, padded to a multiple of 8

}  // namespace

struct syn_heads {
  RawLayer raw[2][9];                // [net][layer]: MLP_for conv1..conv9; MLP_rev conv1..conv5, conv6_1, conv6_2, conv6_3
  bool committed[2] = {false, false};
  // MLP_for: L[0] = conv1 (CUDA cores), 1..4 = conv2..conv5, 5 = conv6 face part, 6 = conv6 point part, 7..9 = conv7..conv9
  GemmLayer lf[10];
  // MLP_rev: 0 = conv1, 1..4 = conv2..conv5, 5 = the three conv6_x heads concatenated (1024 -> 62)
  GemmLayer lr[6];
  // workspace (grown on demand)
  int ws_batch = 0;
  float *bufA = nullptr, *bufB = nullptr, *pf = nullptr, *facevec = nullptr, *addend = nullptr;
  unsigned *rmA = nullptr, *rmB = nullptr, *rm_pf = nullptr, *rm_face = nullptr, *gmax = nullptr;
};

namespace {

void heads_free_ws(syn_heads* s) {
  cudaFree(s->bufA); cudaFree(s->bufB); cudaFree(s->pf); cudaFree(s->facevec); cudaFree(s->addend);
  cudaFree(s->rmA); cudaFree(s->rmB); cudaFree(s->rm_pf); cudaFree(s->rm_face); cudaFree(s->gmax);
  s->bufA = s->bufB = s->pf = s->facevec = s->addend = nullptr;
  s->rmA = s->rmB = s->rm_pf = s->rm_face = s->gmax = nullptr;
  s->ws_batch = 0;
}

int heads_workspace(syn_heads* s, int batch) {
  if (batch <= s->ws_batch) return SYN_OK;
  SYN_CUDA(cudaDeviceSynchronize());
  heads_free_ws(s);
  const size_t M = (size_t)batch * kPts;
  SYN_CUDA(cudaMalloc(&s->bufA, M * 512 * sizeof(float)));
  SYN_CUDA(cudaMalloc(&s->bufB, M * 512 * sizeof(float)));
  SYN_CUDA(cudaMalloc(&s->pf, M * 64 * sizeof(float)));
  SYN_CUDA(cudaMalloc(&s->facevec, (size_t)batch * kFaceVecLd * sizeof(float)));
  SYN_CUDA(cudaMalloc(&s->addend, (size_t)batch * 512 * sizeof(float)));
  SYN_CUDA(cudaMalloc(&s->rmA, M * sizeof(unsigned)));
  SYN_CUDA(cudaMalloc(&s->rmB, M * sizeof(unsigned)));
  SYN_CUDA(cudaMalloc(&s->rm_pf, M * sizeof(unsigned)));
  SYN_CUDA(cudaMalloc(&s->rm_face, (size_t)batch * sizeof(unsigned)));
  SYN_CUDA(cudaMalloc(&s->gmax, (size_t)batch * 1024 * sizeof(unsigned)));
  s->ws_batch = batch;
  return SYN_OK;
}

// conv1 (3 -> 64) on CUDA cores: needs [K][N] fp32 weights
int build_small_k_layer(GemmLayer& L, const std::vector<double>& Wnk, const std::vector<double>& bias, int N, int K) {
  free_layer(L);
  L.N = N; L.K = K; L.act = kActRelu;
  std::vector<float> wkn((size_t)K * N), bs(N);
  for (int n = 0; n < N; ++n) {
    bs[n] = (float)bias[n];
    for (int k = 0; k < K; ++k) wkn[(size_t)k * N + n] = (float)Wnk[(size_t)n * K + k];
  }
  int rc = upload(&L.d_wkn, wkn);
  if (rc != SYN_OK) return rc;
  return upload(&L.d_bias, bs);
}

// conv1 .. conv5 + max-pool over the points, shared by MLP_for and MLP_rev (pointnet_backbone.py:32-38, 91-96).
// Leaves point_features (conv2 output) in s->pf / s->rm_pf and the pooled global features (as fp32 bits) in s->gmax.
int pointnet_trunk(syn_handle* h, syn_heads* s, const GemmLayer* L, const float* lmk, int batch, cudaStream_t st) {
  const int M = batch * kPts;
  small_k_layer_kernel<<<(M + 7) / 8, dim3(32, 8), 0, st>>>(lmk, L[0].d_wkn, L[0].d_bias, s->bufA, s->rmA, M, 3, 64, kPts, kActRelu);
  SYN_LAUNCH_CHECK("small_k_layer_kernel");
  mark(h, st, "pointnet_conv1");
  GemmIO io;
  io.M = M;
  io.A = s->bufA; io.lda = 64; io.rowmax_in = s->rmA; io.out = s->pf; io.rowmax_out = s->rm_pf;
  int rc = launch_gemm(h, L[1], io, st, "pointnet_conv2");
  if (rc != SYN_OK) return rc;
  io.A = s->pf; io.rowmax_in = s->rm_pf; io.out = s->bufA; io.rowmax_out = s->rmA;
  rc = launch_gemm(h, L[2], io, st, "pointnet_conv3");
  if (rc != SYN_OK) return rc;
  io.A = s->bufA; io.rowmax_in = s->rmA; io.out = s->bufB; io.rowmax_out = s->rmB;
  rc = launch_gemm(h, L[3], io, st, "pointnet_conv4");
  if (rc != SYN_OK) return rc;
  SYN_CUDA(cudaMemsetAsync(s->gmax, 0, (size_t)batch * 1024 * sizeof(unsigned), st));
  io.A = s->bufB; io.lda = 128; io.rowmax_in = s->rmB; io.out = nullptr; io.rowmax_out = nullptr;
  io.colmax_out = s->gmax; io.colmax_group = kPts;
  return launch_gemm(h, L[4], io, st, "pointnet_conv5_maxpool");
}

int heads_commit_net(syn_handle* h, int net) {
  syn_heads* s = h->heads;
  const int nl = net == 0 ? 9 : 8;
  for (int i = 0; i < nl; ++i)
    if (!s->raw[net][i].set) return fail(SYN_ERR_STATE, "syn_pointnet_commit: layer %d of net %d not set", i, net);
  std::vector<double> w, b;
  GemmLayer* L = net == 0 ? s->lf : s->lr;
  fold_bn(s->raw[net][0], w, b);
  int rc = build_small_k_layer(L[0], w, b, 64, 3);
  if (rc != SYN_OK) return rc;
  for (int i = 1; i <= 4; ++i) {
    const RawLayer& r = s->raw[net][i];
    fold_bn(r, w, b);
    rc = build_gemm_layer(L[i], w, b, r.cout, r.cin, kActRelu);
    if (rc != SYN_OK) return rc;
  }
  if (net == 0) {
    // conv6 (2418 -> 512): columns [0,64) act on point_features, columns [64,2418) on per-face inputs
    const RawLayer& r = s->raw[0][5];
    fold_bn(r, w, b);
    std::vector<double> wp((size_t)512 * 64), wf((size_t)512 * kFaceVecLd, 0.0), zero(512, 0.0);
    for (int n = 0; n < 512; ++n) {
      for (int k = 0; k < 64; ++k) wp[(size_t)n * 64 + k] = w[(size_t)n * 2418 + k];
      for (int k = 0; k < 2354; ++k) wf[(size_t)n * kFaceVecLd + k] = w[(size_t)n * 2418 + 64 + k];
    }
    rc = build_gemm_layer(L[5], wf, zero, 512, kFaceVecLd, kActNone);
    if (rc != SYN_OK) return rc;
    rc = build_gemm_layer(L[6], wp, b, 512, 64, kActRelu);
    if (rc != SYN_OK) return rc;
    for (int i = 6; i <= 8; ++i) {
      const RawLayer& q = s->raw[0][i];
      fold_bn(q, w, b);
      rc = build_gemm_layer(L[i + 1], w, b, q.cout, q.cin, kActRelu);
      if (rc != SYN_OK) return rc;
    }
  } else {
    // conv6_1 | conv6_2 | conv6_3 (+BN+ReLU each) concatenated: (1024 -> 12 | 40 | 10), pointnet_backbone.py:98-104
    std::vector<double> wc((size_t)62 * 1024), bc(62);
    int n0 = 0;
    for (int i = 5; i <= 7; ++i) {
      const RawLayer& q = s->raw[1][i];
      fold_bn(q, w, b);
      for (int n = 0; n < q.cout; ++n) {
        bc[n0 + n] = b[n];
        for (int k = 0; k < 1024; ++k) wc[(size_t)(n0 + n) * 1024 + k] = w[(size_t)n * 1024 + k];
      }
      n0 += q.cout;
    }
    rc = build_gemm_layer(L[5], wc, bc, 62, 1024, kActRelu);
    if (rc != SYN_OK) return rc;
  }
  s->committed[net] = true;
  return SYN_OK;
}

}  // namespace

void syn_heads_destroy(syn_heads* s) {
  if (s == nullptr) return;
  heads_free_ws(s);
  for (auto& L : s->lf) free_layer(L);
  for (auto& L : s->lr) free_layer(L);
  delete s;
}

extern "C" {

int syn_pointnet_set_layer(syn_handle_t* h, int net, int layer, const float* w_host, int cout, int cin,
                           const float* conv_bias_host, const float* bn_weight_host, const float* bn_bias_host,
                           const float* bn_mean_host, const float* bn_var_host, float eps) {
  if (h == nullptr || w_host == nullptr) return fail(SYN_ERR_INVALID, "syn_pointnet_set_layer: null argument");
  static const int dims_for[9][2] = {{3, 64}, {64, 64}, {64, 64}, {64, 128}, {128, 1024}, {2418, 512}, {512, 256}, {256, 128}, {128, 3}};
  static const int dims_rev[8][2] = {{3, 64}, {64, 64}, {64, 64}, {64, 128}, {128, 1024}, {1024, 12}, {1024, 40}, {1024, 10}};
  const int nl = net == 0 ? 9 : 8;
  if (net < 0 || net > 1 || layer < 0 || layer >= nl) return fail(SYN_ERR_INVALID, "syn_pointnet_set_layer: bad net %d / layer %d", net, layer);
  const int* d = net == 0 ? dims_for[layer] : dims_rev[layer];
  if (cin != d[0] || cout != d[1])
    return fail(SYN_ERR_SHAPE, "syn_pointnet_set_layer: net %d layer %d expects %d -> %d channels, got %d -> %d", net, layer, d[0], d[1], cin, cout);
  if (bn_weight_host == nullptr || bn_bias_host == nullptr || bn_mean_host == nullptr || bn_var_host == nullptr)
    return fail(SYN_ERR_INVALID, "syn_pointnet_set_layer: every PointNet conv is followed by a BatchNorm1d");
  if (h->heads == nullptr) h->heads = new (std::nothrow) syn_heads();
  if (h->heads == nullptr) return fail(SYN_ERR_NOMEM, "syn_pointnet_set_layer: out of host memory");
  RawLayer& r = h->heads->raw[net][layer];
  r.cout = cout; r.cin = cin; r.ksize = 1; r.eps = eps; r.has_bn = true;
  r.w.assign(w_host, w_host + (size_t)cout * cin);
  if (conv_bias_host) r.b.assign(conv_bias_host, conv_bias_host + cout); else r.b.clear();
  r.g.assign(bn_weight_host, bn_weight_host + cout);
  r.beta.assign(bn_bias_host, bn_bias_host + cout);
  r.m.assign(bn_mean_host, bn_mean_host + cout);
  r.v.assign(bn_var_host, bn_var_host + cout);
  r.set = true;
  h->heads->committed[net] = false;
  return SYN_OK;
}

int syn_pointnet_commit(syn_handle_t* h, int net) {
  if (h == nullptr || h->heads == nullptr || net < 0 || net > 1) return fail(SYN_ERR_STATE, "syn_pointnet_commit: no layers set");
  DeviceGuard g(h->device);
  if (h->d_err == nullptr) return fail(SYN_ERR_STATE, "syn_pointnet_commit: commit the backbone first (syn_commit)");
  return heads_commit_net(h, net);
}

int syn_mlp_for(syn_handle_t* h, const float* lmk_dev, const float* pool1280_dev, const float* params62_dev, int batch,
                float* residual_dev, float* refined_dev, void* stream) {
  SYN_CHECK_READY(h, "syn_mlp_for");
  if (h->heads == nullptr || !h->heads->committed[0]) return fail(SYN_ERR_STATE, "syn_mlp_for: MLP_for weights not committed");
  if (lmk_dev == nullptr || pool1280_dev == nullptr || params62_dev == nullptr || batch <= 0 ||
      (residual_dev == nullptr && refined_dev == nullptr))
    return fail(SYN_ERR_INVALID, "syn_mlp_for: bad argument");
  DeviceGuard g(h->device);
  syn_heads* s = h->heads;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = heads_workspace(s, batch);
  if (rc != SYN_OK) return rc;
  if (h->timing) { mark(h, st, "start"); h->launches--; }
  rc = pointnet_trunk(h, s, s->lf, lmk_dev, batch, st);
  if (rc != SYN_OK) return rc;
  const int M = batch * kPts;
  pointnet_face_vector_kernel<<<batch, 256, 0, st>>>(s->gmax, pool1280_dev, params62_dev, s->facevec, batch, kFaceVecLd);
  SYN_LAUNCH_CHECK("pointnet_face_vector_kernel");
  mark(h, st, "pointnet_face_vector");
  rowmax_kernel<<<(batch + 7) / 8, dim3(32, 8), 0, st>>>(s->facevec, s->rm_face, batch, kFaceVecLd, kFaceVecLd);
  SYN_LAUNCH_CHECK("rowmax_kernel");
  mark(h, st, "pointnet_face_rowmax");
  GemmIO io;
  io.M = batch; io.A = s->facevec; io.lda = kFaceVecLd; io.rowmax_in = s->rm_face; io.out = s->addend;
  rc = launch_gemm(h, s->lf[5], io, st, "pointnet_conv6_face");
  if (rc != SYN_OK) return rc;
  io = GemmIO();
  io.M = M; io.A = s->pf; io.lda = 64; io.rowmax_in = s->rm_pf; io.out = s->bufA; io.rowmax_out = s->rmA;
  io.addend = s->addend; io.addend_group = kPts;
  rc = launch_gemm(h, s->lf[6], io, st, "pointnet_conv6_point");
  if (rc != SYN_OK) return rc;
  io = GemmIO();
  io.M = M; io.A = s->bufA; io.lda = 512; io.rowmax_in = s->rmA; io.out = s->bufB; io.rowmax_out = s->rmB;
  rc = launch_gemm(h, s->lf[7], io, st, "pointnet_conv7");
  if (rc != SYN_OK) return rc;
  io.A = s->bufB; io.lda = 256; io.rowmax_in = s->rmB; io.out = s->bufA; io.rowmax_out = s->rmA;
  rc = launch_gemm(h, s->lf[8], io, st, "pointnet_conv8");
  if (rc != SYN_OK) return rc;
  io.A = s->bufA; io.lda = 128; io.rowmax_in = s->rmA; io.out = s->bufB; io.rowmax_out = nullptr;
  rc = launch_gemm(h, s->lf[9], io, st, "pointnet_conv9");
  if (rc != SYN_OK) return rc;
  const int n = batch * 3 * kPts;
  pointnet_residual_kernel<<<(n + 255) / 256, 256, 0, st>>>(s->bufB, 3, lmk_dev, residual_dev, refined_dev, batch, kPts);
  SYN_LAUNCH_CHECK("pointnet_residual_kernel");
  mark(h, st, "pointnet_residual");
  return SYN_OK;
}

int syn_mlp_rev(syn_handle_t* h, const float* lmk_dev, int batch, float* params62_dev, void* stream) {
  SYN_CHECK_READY(h, "syn_mlp_rev");
  if (h->heads == nullptr || !h->heads->committed[1]) return fail(SYN_ERR_STATE, "syn_mlp_rev: MLP_rev weights not committed");
  if (lmk_dev == nullptr || params62_dev == nullptr || batch <= 0) return fail(SYN_ERR_INVALID, "syn_mlp_rev: bad argument");
  DeviceGuard g(h->device);
  syn_heads* s = h->heads;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = heads_workspace(s, batch);
  if (rc != SYN_OK) return rc;
  if (h->timing) { mark(h, st, "start"); h->launches--; }
  rc = pointnet_trunk(h, s, s->lr, lmk_dev, batch, st);
  if (rc != SYN_OK) return rc;
  // global features (B,1024) as floats, their row maxima, then the three heads as one GEMM
  const size_t n = (size_t)batch * 1024;
  bits_to_float_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(s->gmax, s->bufA, n);
  SYN_LAUNCH_CHECK("bits_to_float_kernel");
  mark(h, st, "pointnet_global_features");
  rowmax_kernel<<<(batch + 7) / 8, dim3(32, 8), 0, st>>>(s->bufA, s->rm_face, batch, 1024, 1024);
  SYN_LAUNCH_CHECK("rowmax_kernel");
  mark(h, st, "pointnet_global_rowmax");
  GemmIO io;
  io.M = batch; io.A = s->bufA; io.lda = 1024; io.rowmax_in = s->rm_face; io.out = params62_dev;
  return launch_gemm(h, s->lr[5], io, st, "pointnet_rev_heads");
}

// Debug only: copy one workspace buffer of the PointNet heads to the host after a call (0 = point_features (B*68,64),
// 1 = face vector (B,2360), 2 = conv6 face part (B,512), 3 = bufA, 4 = bufB, 5 = max-pooled conv5 as floats (B,1024)).
int syn_debug_heads_buffer(syn_handle_t* h, int which, float* out_host, int64_t n) {
  if (h == nullptr || h->heads == nullptr || out_host == nullptr || n <= 0) return fail(SYN_ERR_INVALID, "syn_debug_heads_buffer: bad argument");
  DeviceGuard g(h->device);
  SYN_CUDA(cudaDeviceSynchronize());
  syn_heads* s = h->heads;
  const void* src = which == 0 ? (const void*)s->pf : which == 1 ? (const void*)s->facevec : which == 2 ? (const void*)s->addend
                    : which == 3 ? (const void*)s->bufA : which == 4 ? (const void*)s->bufB : (const void*)s->gmax;
  if (src == nullptr) return fail(SYN_ERR_STATE, "syn_debug_heads_buffer: workspace not allocated");
  SYN_CUDA(cudaMemcpy(out_host, src, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
  return SYN_OK;
}

int syn_wing_loss(syn_handle_t* h, const float* pred_dev, const float* target_dev, int batch, int n_pts, float* out_dev,
                  void* stream) {
  if (h == nullptr || pred_dev == nullptr || target_dev == nullptr || out_dev == nullptr || batch <= 0 || n_pts <= 0)
    return fail(SYN_ERR_INVALID, "syn_wing_loss: bad argument");
  DeviceGuard g(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  wing_loss_kernel<<<1, 1024, 0, st>>>(pred_dev, target_dev, (size_t)batch * 3 * n_pts, 10.0f, 2.0f, out_dev);
  SYN_LAUNCH_CHECK("wing_loss_kernel");
  mark(h, st, "wing_loss_kernel");
  return SYN_OK;
}

int syn_param_loss(syn_handle_t* h, const float* input_dev, const float* target_dev, int batch, int mode, float* out_dev,
                   void* stream) {
  if (h == nullptr || input_dev == nullptr || target_dev == nullptr || out_dev == nullptr || batch <= 0 || mode < 0 || mode > 1)
    return fail(SYN_ERR_INVALID, "syn_param_loss: bad argument");
  DeviceGuard g(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  param_loss_kernel<<<(batch + 127) / 128, 128, 0, st>>>(input_dev, target_dev, batch, mode, out_dev);
  SYN_LAUNCH_CHECK("param_loss_kernel");
  mark(h, st, "param_loss_kernel");
  return SYN_OK;
}

}  // extern "C"
