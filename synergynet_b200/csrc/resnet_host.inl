// Host side of the ResNet-50 backbone variant (included by synergy_b200.cu after heads_host.inl):
// reference backbone_nets/resnet_backbone.py:120-249 (Bottleneck, ResNet._forward_impl), BASELINE.json configs[4].
// 53 convolutions in execution order -- index 0 = conv1 (7x7/s2), then per bottleneck conv1 (1x1), conv2 (3x3, carries
// the stride), conv3 (1x1) and, for the first block of every stage, the 1x1 downsample of the shortcut -- plus the four
// Linear heads concatenated in the reference's output order ori | shape | exp | tex (:242-246).

namespace {

struct RsConv { int cin, cout, ksize, stride, hin, hout, stage, block, role; };   // role: 0 stem, 1..3 conv1..3, 4 downsample

inline std::vector<RsConv> resnet50_plan() {
  std::vector<RsConv> v;
  v.push_back({3, 64, 7, 2, 120, 60, 0, 0, 0});
  static const int planes[4] = {64, 128, 256, 512}, nblk[4] = {3, 4, 6, 3}, strd[4] = {1, 2, 2, 2};
  int cin = 64, h = 30;
  for (int s = 0; s < 4; ++s)
    for (int j = 0; j < nblk[s]; ++j) {
      const int st = (j == 0) ? strd[s] : 1, p = planes[s];
      const int ho = (h + 2 - 3) / st + 1;
      v.push_back({cin, p, 1, 1, h, h, s, j, 1});
      v.push_back({p, p, 3, st, h, ho, s, j, 2});
      v.push_back({p, 4 * p, 1, 1, ho, ho, s, j, 3});
      if (j == 0) v.push_back({cin, 4 * p, 1, st, h, ho, s, j, 4});
      cin = 4 * p;
      h = ho;
    }
  return v;
}

}  // namespace

struct syn_resnet {
  std::vector<RsConv> plan = resnet50_plan();
  std::vector<RawLayer> raw = std::vector<RawLayer>(53);
  RawLayer fc;                       // (102, 2048) + bias
  std::vector<GemmLayer> L = std::vector<GemmLayer>(54);   // 0: stem (d_wkn), 1..52 convs, 53: heads
  bool committed = false;
  int ws_batch = 0;
  float *bx = nullptr, *by = nullptr, *t1 = nullptr, *t2 = nullptr, *ds = nullptr, *stem = nullptr;
  unsigned *rx = nullptr, *ry = nullptr, *r1 = nullptr, *r2 = nullptr, *rstem = nullptr;
};

namespace {

void resnet_free_ws(syn_resnet* s) {
  cudaFree(s->bx); cudaFree(s->by); cudaFree(s->t1); cudaFree(s->t2); cudaFree(s->ds); cudaFree(s->stem);
  cudaFree(s->rx); cudaFree(s->ry); cudaFree(s->r1); cudaFree(s->r2); cudaFree(s->rstem);
  s->bx = s->by = s->t1 = s->t2 = s->ds = s->stem = nullptr;
  s->rx = s->ry = s->r1 = s->r2 = s->rstem = nullptr;
  s->ws_batch = 0;
}

int resnet_workspace(syn_resnet* s, int batch) {
  if (batch <= s->ws_batch) return SYN_OK;
  SYN_CUDA(cudaDeviceSynchronize());
  resnet_free_ws(s);
  const size_t b = (size_t)batch;
  SYN_CUDA(cudaMalloc(&s->stem, b * 60 * 60 * 64 * sizeof(float)));
  SYN_CUDA(cudaMalloc(&s->bx, b * 30 * 30 * 256 * sizeof(float)));      // block in / out (largest: 30x30x256)
  SYN_CUDA(cudaMalloc(&s->by, b * 30 * 30 * 256 * sizeof(float)));
  SYN_CUDA(cudaMalloc(&s->t1, b * 30 * 30 * 128 * sizeof(float)));      // conv1 output (largest: stage 2 block 0)
  SYN_CUDA(cudaMalloc(&s->t2, b * 30 * 30 * 64 * sizeof(float)));       // conv2 output
  SYN_CUDA(cudaMalloc(&s->ds, b * 30 * 30 * 256 * sizeof(float)));      // downsampled shortcut
  SYN_CUDA(cudaMalloc(&s->rstem, b * 3600 * sizeof(unsigned)));
  SYN_CUDA(cudaMalloc(&s->rx, b * 900 * sizeof(unsigned)));
  SYN_CUDA(cudaMalloc(&s->ry, b * 900 * sizeof(unsigned)));
  SYN_CUDA(cudaMalloc(&s->r1, b * 900 * sizeof(unsigned)));
  SYN_CUDA(cudaMalloc(&s->r2, b * 900 * sizeof(unsigned)));
  s->ws_batch = batch;
  return SYN_OK;
}

}  // namespace

void syn_resnet_destroy(syn_resnet* s) {
  if (s == nullptr) return;
  resnet_free_ws(s);
  for (auto& l : s->L) free_layer(l);
  delete s;
}

extern "C" {

int syn_resnet_num_convs(void) { return 53; }

int syn_resnet_conv_desc(int idx, syn_conv_desc_t* out) {
  static const std::vector<RsConv> plan = resnet50_plan();
  if (idx < 0 || idx >= (int)plan.size() || out == nullptr) return fail(SYN_ERR_INVALID, "syn_resnet_conv_desc: bad index %d", idx);
  const RsConv& c = plan[idx];
  out->cin = c.cin; out->cout = c.cout; out->ksize = c.ksize; out->stride = c.stride; out->groups = 1;
  out->relu6 = 0; out->h_in = c.hin; out->h_out = c.hout; out->residual = c.role == 3;
  return SYN_OK;
}

int syn_resnet_set_conv(syn_handle_t* h, int idx, const float* w_host, int64_t w_numel, const float* bn_weight_host,
                        const float* bn_bias_host, const float* bn_mean_host, const float* bn_var_host, float eps) {
  if (h == nullptr || w_host == nullptr || bn_weight_host == nullptr || bn_bias_host == nullptr || bn_mean_host == nullptr ||
      bn_var_host == nullptr)
    return fail(SYN_ERR_INVALID, "syn_resnet_set_conv: null argument");
  if (h->resnet == nullptr) h->resnet = new (std::nothrow) syn_resnet();
  if (h->resnet == nullptr) return fail(SYN_ERR_NOMEM, "syn_resnet_set_conv: out of host memory");
  syn_resnet* s = h->resnet;
  if (idx < 0 || idx >= 53) return fail(SYN_ERR_INVALID, "syn_resnet_set_conv: bad index %d", idx);
  const RsConv& c = s->plan[idx];
  const int64_t want = (int64_t)c.cout * c.cin * c.ksize * c.ksize;
  if (w_numel != want) return fail(SYN_ERR_SHAPE, "syn_resnet_set_conv: conv %d expects %lld weights, got %lld", idx, (long long)want, (long long)w_numel);
  RawLayer& r = s->raw[idx];
  r.cout = c.cout; r.cin = c.cin; r.ksize = c.ksize; r.eps = eps; r.has_bn = true;
  r.w.assign(w_host, w_host + want);
  r.b.clear();                                               // resnet convs have bias=False (:41-49)
  r.g.assign(bn_weight_host, bn_weight_host + c.cout);
  r.beta.assign(bn_bias_host, bn_bias_host + c.cout);
  r.m.assign(bn_mean_host, bn_mean_host + c.cout);
  r.v.assign(bn_var_host, bn_var_host + c.cout);
  r.set = true;
  s->committed = false;
  return SYN_OK;
}

int syn_resnet_set_heads(syn_handle_t* h, const float* w_host, const float* b_host) {
  if (h == nullptr || w_host == nullptr || b_host == nullptr) return fail(SYN_ERR_INVALID, "syn_resnet_set_heads: null argument");
  if (h->resnet == nullptr) h->resnet = new (std::nothrow) syn_resnet();
  if (h->resnet == nullptr) return fail(SYN_ERR_NOMEM, "syn_resnet_set_heads: out of host memory");
  RawLayer& r = h->resnet->fc;
  r.cout = 102; r.cin = 2048; r.ksize = 1; r.has_bn = false;
  r.w.assign(w_host, w_host + (size_t)102 * 2048);
  r.b.assign(b_host, b_host + 102);
  r.set = true;
  h->resnet->committed = false;
  return SYN_OK;
}

int syn_resnet_commit(syn_handle_t* h) {
  if (h == nullptr || h->resnet == nullptr) return fail(SYN_ERR_STATE, "syn_resnet_commit: no layers set");
  if (h->d_err == nullptr) return fail(SYN_ERR_STATE, "syn_resnet_commit: commit the MobileNetV2 path first (syn_commit allocates the shared state)");
  DeviceGuard g(h->device);
  syn_resnet* s = h->resnet;
  for (int i = 0; i < 53; ++i)
    if (!s->raw[i].set) return fail(SYN_ERR_STATE, "syn_resnet_commit: conv %d not set", i);
  if (!s->fc.set) return fail(SYN_ERR_STATE, "syn_resnet_commit: heads not set");
  std::vector<double> w, b, wk;
  {  // stem: [147][64] with k = (ci*7+ky)*7+kx, the OIHW order of the reference weight
    fold_bn(s->raw[0], w, b);
    int rc = build_small_k_layer(s->L[0], w, b, 64, 147);
    if (rc != SYN_OK) return rc;
  }
  for (int i = 1; i < 53; ++i) {
    const RsConv& c = s->plan[i];
    fold_bn(s->raw[i], w, b);
    const int K = c.ksize * c.ksize * c.cin;
    wk.assign((size_t)c.cout * K, 0.0);                     // GEMM k order: (ky*ks + kx)*C + c  <-  OIHW [n][c][ky][kx]
    for (int n = 0; n < c.cout; ++n)
      for (int ci = 0; ci < c.cin; ++ci)
        for (int t = 0; t < c.ksize * c.ksize; ++t)
          wk[(size_t)n * K + (size_t)t * c.cin + ci] = w[((size_t)n * c.cin + ci) * c.ksize * c.ksize + t];
    const int act = (c.role == 4) ? kActNone : kActRelu;     // conv3's ReLU comes after the shortcut add (:140-144)
    int rc = build_gemm_layer(s->L[i], wk, b, c.cout, K, act);
    if (rc != SYN_OK) return rc;
    if (c.ksize > 1 || c.stride > 1) {                       // implicit GEMM over the k x k x C patch (or the strided pixel grid)
      s->L[i].ksize = c.ksize; s->L[i].stride = c.stride; s->L[i].pad = c.ksize / 2;
    }                                                        // else: plain rows, pixel m of the input = row m of the GEMM
  }
  fold_bn(s->fc, w, b);
  int rc = build_gemm_layer(s->L[53], w, b, 102, 2048, kActNone);
  if (rc != SYN_OK) return rc;
  s->committed = true;
  return SYN_OK;
}

int syn_resnet50_forward(syn_handle_t* h, const float* x_dev, int batch, float* out102_dev, float* pool2048_dev, void* stream) {
  SYN_CHECK_READY(h, "syn_resnet50_forward");
  if (h->resnet == nullptr || !h->resnet->committed) return fail(SYN_ERR_STATE, "syn_resnet50_forward: ResNet-50 weights not committed");
  if (x_dev == nullptr || out102_dev == nullptr || batch <= 0) return fail(SYN_ERR_INVALID, "syn_resnet50_forward: bad argument");
  DeviceGuard g(h->device);
  syn_resnet* s = h->resnet;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = resnet_workspace(s, batch);
  if (rc != SYN_OK) return rc;
  if (h->timing) { mark(h, st, "start"); h->launches--; }
  resnet_stem_kernel<<<batch * 60, kRsStemThreads, 0, st>>>(x_dev, s->L[0].d_wkn, s->L[0].d_bias, s->stem, s->rstem, batch);
  SYN_LAUNCH_CHECK("resnet_stem_kernel");
  mark(h, st, "resnet_stem_kernel");
  {
    const int npix = batch * 900;
    maxpool3x3s2_kernel<<<(npix + 7) / 8, dim3(32, 8), 0, st>>>(s->stem, s->bx, s->rx, batch, 60, 30, 64);
    SYN_LAUNCH_CHECK("maxpool3x3s2_kernel");
    mark(h, st, "maxpool3x3s2_kernel");
  }
  float *X = s->bx, *Y = s->by;
  unsigned *RX = s->rx, *RY = s->ry;
  auto conv = [&](int idx, const float* in, const unsigned* rin, float* out, unsigned* rout, const float* residual) -> int {
    const RsConv& c = s->plan[idx];
    GemmIO io;
    io.M = batch * c.hout * c.hout;
    io.A = in; io.lda = c.cin; io.rowmax_in = rin; io.out = out; io.rowmax_out = rout; io.residual = residual;
    io.H = c.hin; io.W = c.hin; io.C = c.cin; io.HO = c.hout; io.WO = c.hout;
    static const char* const names[5] = {"", "resnet_conv1x1_a", "resnet_conv3x3", "resnet_conv1x1_b", "resnet_downsample"};
    return launch_gemm(h, s->L[idx], io, st, names[c.role]);
  };
  int idx = 1;
  while (idx < 53) {
    const bool has_ds = (idx + 3 < 53 + 1) && s->plan[idx].block == 0;
    rc = conv(idx, X, RX, s->t1, s->r1, nullptr);                          // conv1 + bn1 + relu (:126-128)
    if (rc != SYN_OK) return rc;
    rc = conv(idx + 1, s->t1, s->r1, s->t2, s->r2, nullptr);               // conv2 + bn2 + relu (:130-132)
    if (rc != SYN_OK) return rc;
    const float* identity = X;
    if (has_ds) {
      rc = conv(idx + 3, X, RX, s->ds, nullptr, nullptr);                  // downsample(x) (:137-138)
      if (rc != SYN_OK) return rc;
      identity = s->ds;
    }
    rc = conv(idx + 2, s->t2, s->r2, Y, RY, identity);                     // conv3 + bn3, += identity, relu (:134-142)
    if (rc != SYN_OK) return rc;
    std::swap(X, Y);
    std::swap(RX, RY);
    idx += has_ds ? 4 : 3;
  }
  // avgpool + flatten (:236-237), then the four heads (:239-246)
  float* pooled = pool2048_dev ? pool2048_dev : s->t1;
  avgpool_kernel<<<batch, 256, 0, st>>>(X, pooled, s->r1, 16, 2048);
  SYN_LAUNCH_CHECK("avgpool_kernel");
  mark(h, st, "avgpool_kernel");
  GemmIO io;
  io.M = batch; io.A = pooled; io.lda = 2048; io.rowmax_in = s->r1; io.out = out102_dev;
  return launch_gemm(h, s->L[53], io, st, "resnet_heads");
}

}  // extern "C"
