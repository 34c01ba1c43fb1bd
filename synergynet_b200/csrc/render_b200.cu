// C ABI of the stages either side of the 3DMM path: Sim3DR normals / lighting /
// rasterisation of the dense meshes, and the FaceBoxes box decode + greedy NMS that produces the crops.
// Handle-free: device pointers and workspaces belong to the caller (include/synergy_b200.h states the sizes).
#include "kernels_render.cuh"
#include "kernels_detect.cuh"

#include <cmath>
#include <new>
#include <vector>

using namespace syn;

namespace {

int check_mesh(const float* v, long long sb, int sv, int sc, int batch, int nver, MeshView& m) {
  if (!v || batch <= 0 || nver <= 0 || sv <= 0 || sc <= 0 || (batch > 1 && sb <= 0))
    return fail(SYN_ERR_INVALID, "mesh view: null pointer, empty batch or non-positive stride");
  m.v = v; m.sb = sb; m.sv = sv; m.sc = sc; m.nver = nver; m.batch = batch;
  return SYN_OK;
}

}  // namespace

extern "C" {

int syn_mesh_incidence_host(const int32_t* tri_host, int ntri, int nver, int32_t* start_out, int32_t* list_out) {
  if (!tri_host || !start_out || !list_out || ntri < 0 || nver <= 0) return fail(SYN_ERR_INVALID, "syn_mesh_incidence_host: bad argument");
  for (int i = 0; i < 3 * ntri; ++i)
    if (tri_host[i] < 0 || tri_host[i] >= nver) return fail(SYN_ERR_SHAPE, "triangle %d references vertex %d of %d", i / 3, tri_host[i], nver);
  for (int v = 0; v <= nver; ++v) start_out[v] = 0;
  for (int i = 0; i < 3 * ntri; ++i) ++start_out[tri_host[i] + 1];
  for (int v = 0; v < nver; ++v) start_out[v + 1] += start_out[v];
  std::vector<int32_t> fill(start_out, start_out + nver);
  for (int t = 0; t < ntri; ++t)                 // triangles in order: every vertex's list comes out ascending
    for (int k = 0; k < 3; ++k) list_out[fill[tri_host[3 * t + k]]++] = t;
  return SYN_OK;
}

int syn_mesh_normals(const float* vertices_dev, int64_t stride_mesh, int stride_vertex, int stride_coord, int batch, int nver,
                     const int32_t* tri_dev, int ntri, const int32_t* inc_start_dev, const int32_t* inc_tri_dev,
                     float* tri_normals_ws_dev, float* normals_dev, void* stream) {
  MeshView m;
  if (int rc = check_mesh(vertices_dev, stride_mesh, stride_vertex, stride_coord, batch, nver, m)) return rc;
  if (!tri_dev || ntri <= 0 || !inc_start_dev || !inc_tri_dev || !tri_normals_ws_dev || !normals_dev)
    return fail(SYN_ERR_INVALID, "syn_mesh_normals: null pointer or no triangles");
  cudaStream_t st = (cudaStream_t)stream;
  tri_normal_kernel<<<dim3((ntri + 255) / 256, batch), 256, 0, st>>>(m, tri_dev, ntri, tri_normals_ws_dev);
  SYN_LAUNCH_CHECK("tri_normal_kernel");
  vertex_normal_kernel<<<dim3((nver + 255) / 256, batch), 256, 0, st>>>(nver, ntri, tri_normals_ws_dev, inc_start_dev, inc_tri_dev, normals_dev);
  SYN_LAUNCH_CHECK("vertex_normal_kernel");
  return SYN_OK;
}

int syn_mesh_lighting(const float* vertices_dev, int64_t stride_mesh, int stride_vertex, int stride_coord, int batch, int nver,
                      const float* normals_dev, const syn_light_cfg_t* cfg, const float* texture_dev, uint32_t* stats_ws_dev,
                      float* colors_dev, void* stream) {
  MeshView m;
  if (int rc = check_mesh(vertices_dev, stride_mesh, stride_vertex, stride_coord, batch, nver, m)) return rc;
  if (!normals_dev || !cfg || !stats_ws_dev || !colors_dev) return fail(SYN_ERR_INVALID, "syn_mesh_lighting: null pointer");
  rmath::LightCfg c;
  c.intensity_ambient = cfg->intensity_ambient;
  c.intensity_directional = cfg->intensity_directional;
  c.intensity_specular = cfg->intensity_specular;
  c.specular_exp = cfg->specular_exp;
  for (int k = 0; k < 3; ++k) {
    c.color_ambient[k] = cfg->color_ambient[k];
    c.color_directional[k] = cfg->color_directional[k];
    c.light_pos[k] = cfg->light_pos[k];
    c.view_pos[k] = cfg->view_pos[k];
  }
  cudaStream_t st = (cudaStream_t)stream;
  SYN_CUDA(cudaMemsetAsync(stats_ws_dev, 0, sizeof(uint32_t) * 6 * batch, st));
  const int blocks = min((nver + 255) / 256, 64);
  mesh_extent_kernel<<<dim3(blocks, batch), 256, 0, st>>>(m, stats_ws_dev);
  SYN_LAUNCH_CHECK("mesh_extent_kernel");
  vertex_light_kernel<<<dim3((nver + 255) / 256, batch), 256, 0, st>>>(m, normals_dev, stats_ws_dev, c, texture_dev, colors_dev);
  SYN_LAUNCH_CHECK("vertex_light_kernel");
  return SYN_OK;
}

int syn_rasterize(uint8_t* image_dev, int height, int width, int channels, const float* vertices_dev, int64_t stride_mesh,
                  int stride_vertex, int stride_coord, int batch, int nver, const int32_t* tri_dev, int ntri,
                  const float* colors_dev, float alpha, int reverse, uint64_t* keys_ws_dev, float* depth_out_dev, void* stream) {
  MeshView m;
  if (int rc = check_mesh(vertices_dev, stride_mesh, stride_vertex, stride_coord, batch, nver, m)) return rc;
  if (!image_dev || !tri_dev || !colors_dev || !keys_ws_dev || height <= 0 || width <= 0 || channels <= 0 || ntri < 0)
    return fail(SYN_ERR_INVALID, "syn_rasterize: null pointer or empty image");
  if (alpha != 1.0f)
    return fail(SYN_ERR_UNSUPPORTED, "syn_rasterize: alpha = %g; only alpha = 1 (the value Sim3DR.rasterize always passes) has an "
                                     "order-free result", (double)alpha);
  cudaStream_t st = (cudaStream_t)stream;
  SYN_CUDA(cudaMemsetAsync(keys_ws_dev, 0, sizeof(uint64_t) * (size_t)batch * height * width, st));
  if (ntri > 0) {
    raster_depth_kernel<<<dim3((ntri + 255) / 256, batch), 256, 0, st>>>(m, tri_dev, ntri, width, height,
                                                                         reinterpret_cast<unsigned long long*>(keys_ws_dev));
    SYN_LAUNCH_CHECK("raster_depth_kernel");
  }
  raster_resolve_kernel<<<dim3((width + 31) / 32, (height + 7) / 8), dim3(32, 8), 0, st>>>(
      m, tri_dev, colors_dev, channels, width, height, alpha, reverse, reinterpret_cast<const unsigned long long*>(keys_ws_dev),
      image_dev, depth_out_dev);
  SYN_LAUNCH_CHECK("raster_resolve_kernel");
  return SYN_OK;
}

int syn_nms(const float* dets_dev, int n, double thresh, int mode, uint64_t* mask_ws_dev, int32_t* keep_dev, int32_t* n_keep_dev,
            void* stream) {
  if (n < 0 || !n_keep_dev || (n > 0 && (!dets_dev || !mask_ws_dev || !keep_dev))) return fail(SYN_ERR_INVALID, "syn_nms: bad argument");
  if (mode != SYN_NMS_CPU_NMS && mode != SYN_NMS_PY_CPU_NMS) return fail(SYN_ERR_INVALID, "syn_nms: unknown mode %d", mode);
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {                                   // nms_wrapper.py:16-17: no detections, empty keep list
    SYN_CUDA(cudaMemsetAsync(n_keep_dev, 0, sizeof(int32_t), st));
    return SYN_OK;
  }
  const int words = (n + 63) / 64;
  if ((size_t)words * 8 > 200 * 1024) return fail(SYN_ERR_SHAPE, "syn_nms: %d boxes exceed the scan kernel's shared memory", n);
  nms_mask_kernel<<<dim3((words + 31) / 32, (n + 7) / 8), dim3(32, 8), 0, st>>>(dets_dev, n, thresh, mode == SYN_NMS_CPU_NMS ? 1 : 0,
                                                                               reinterpret_cast<unsigned long long*>(mask_ws_dev));
  SYN_LAUNCH_CHECK("nms_mask_kernel");
  if (words * 8 > 48 * 1024)
    SYN_CUDA(cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, words * 8));
  nms_scan_kernel<<<1, kNmsScanThreads, words * 8, st>>>(reinterpret_cast<const unsigned long long*>(mask_ws_dev), n, keep_dev, n_keep_dev);
  SYN_LAUNCH_CHECK("nms_scan_kernel");
  return SYN_OK;
}

int syn_faceboxes_num_priors(int im_height, int im_width) {
  if (im_height <= 0 || im_width <= 0) return -1;
  return faceboxes_num_priors(im_height, im_width);
}

int syn_faceboxes_decode(const float* loc_dev, const float* conf_dev, int im_height, int im_width, float box_scale_w,
                         float box_scale_h, float scale, float conf_thresh, int top_k, int32_t* cand_ws_dev, float* dets_dev,
                         int32_t* n_dets_dev, void* stream) {
  if (!loc_dev || !conf_dev || !cand_ws_dev || !dets_dev || !n_dets_dev || im_height <= 0 || im_width <= 0 || top_k <= 0 || !(scale > 0.f))
    return fail(SYN_ERR_INVALID, "syn_faceboxes_decode: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int np = faceboxes_num_priors(im_height, im_width);
  SYN_CUDA(cudaMemsetAsync(n_dets_dev, 0, sizeof(int32_t), st));
  SYN_CUDA(cudaMemsetAsync(cand_ws_dev, 0, sizeof(int32_t), st));
  faceboxes_select_kernel<<<(np + 255) / 256, 256, 0, st>>>(conf_dev, np, conf_thresh, cand_ws_dev);
  SYN_LAUNCH_CHECK("faceboxes_select_kernel");
  faceboxes_rank_decode_kernel<<<(np + 127) / 128, 128, 0, st>>>(loc_dev, conf_dev, im_height, im_width, box_scale_w, box_scale_h, scale,
                                                                top_k, cand_ws_dev, dets_dev, n_dets_dev);
  SYN_LAUNCH_CHECK("faceboxes_rank_decode_kernel");
  return SYN_OK;
}

}  // extern "C"

// ---- the detector network: host side (FaceBoxes/models/faceboxes.py:68-150) ------------------------------------------------
namespace {

struct FbLayer { const char* name; int cin, cout, k, stride, pad, bn, act; };
// execution order; inception layers are 2 + 7 * block + {0 branch1x1, 1 branch1x1_2, 2 branch3x3_reduce, 3 branch3x3,
// 4 branch3x3_reduce_2, 5 branch3x3_2, 6 branch3x3_3} (faceboxes.py:21-47)
const FbLayer kFbLayers[33] = {
    {"conv1", 3, 24, 7, 4, 3, 1, 2},         {"conv2", 48, 64, 5, 2, 2, 1, 2},
    {"inception1.branch1x1", 128, 32, 1, 1, 0, 1, 1},       {"inception1.branch1x1_2", 128, 32, 1, 1, 0, 1, 1},
    {"inception1.branch3x3_reduce", 128, 24, 1, 1, 0, 1, 1}, {"inception1.branch3x3", 24, 32, 3, 1, 1, 1, 1},
    {"inception1.branch3x3_reduce_2", 128, 24, 1, 1, 0, 1, 1}, {"inception1.branch3x3_2", 24, 32, 3, 1, 1, 1, 1},
    {"inception1.branch3x3_3", 32, 32, 3, 1, 1, 1, 1},
    {"inception2.branch1x1", 128, 32, 1, 1, 0, 1, 1},       {"inception2.branch1x1_2", 128, 32, 1, 1, 0, 1, 1},
    {"inception2.branch3x3_reduce", 128, 24, 1, 1, 0, 1, 1}, {"inception2.branch3x3", 24, 32, 3, 1, 1, 1, 1},
    {"inception2.branch3x3_reduce_2", 128, 24, 1, 1, 0, 1, 1}, {"inception2.branch3x3_2", 24, 32, 3, 1, 1, 1, 1},
    {"inception2.branch3x3_3", 32, 32, 3, 1, 1, 1, 1},
    {"inception3.branch1x1", 128, 32, 1, 1, 0, 1, 1},       {"inception3.branch1x1_2", 128, 32, 1, 1, 0, 1, 1},
    {"inception3.branch3x3_reduce", 128, 24, 1, 1, 0, 1, 1}, {"inception3.branch3x3", 24, 32, 3, 1, 1, 1, 1},
    {"inception3.branch3x3_reduce_2", 128, 24, 1, 1, 0, 1, 1}, {"inception3.branch3x3_2", 24, 32, 3, 1, 1, 1, 1},
    {"inception3.branch3x3_3", 32, 32, 3, 1, 1, 1, 1},
    {"conv3_1", 128, 128, 1, 1, 0, 1, 1},    {"conv3_2", 128, 256, 3, 2, 1, 1, 1},
    {"conv4_1", 256, 128, 1, 1, 0, 1, 1},    {"conv4_2", 128, 256, 3, 2, 1, 1, 1},
    {"loc.0", 128, 84, 3, 1, 1, 0, 0},       {"loc.1", 256, 4, 3, 1, 1, 0, 0},       {"loc.2", 256, 4, 3, 1, 1, 0, 0},
    {"conf.0", 128, 42, 3, 1, 1, 0, 0},      {"conf.1", 256, 2, 3, 1, 1, 0, 0},      {"conf.2", 256, 2, 3, 1, 1, 0, 0},
};

inline int conv_out(int n, int k, int s, int p) { return (n + 2 * p - k) / s + 1; }

}  // namespace

struct syn_fb {
  int device = 0;
  std::vector<float> w[33], b[33];          // folded [K][cout] weights and bias, host
  bool set[33] = {};
  float* d_w[33] = {};
  float* d_b[33] = {};
  bool committed = false;
  // workspace for the current image size
  int ws_h = 0, ws_w = 0;
  float *c1 = nullptr, *p1 = nullptr, *c2 = nullptr, *xa = nullptr, *xb = nullptr, *avg = nullptr, *r1 = nullptr, *r2 = nullptr,
        *t3 = nullptr, *c31 = nullptr, *c32 = nullptr, *c41 = nullptr, *c42 = nullptr;
  int64_t launches = 0;
};

namespace {

void fb_free_ws(syn_fb* f) {
  float** bufs[] = {&f->c1, &f->p1, &f->c2, &f->xa, &f->xb, &f->avg, &f->r1, &f->r2, &f->t3, &f->c31, &f->c32, &f->c41, &f->c42};
  for (float** q : bufs) { cudaFree(*q); *q = nullptr; }
  f->ws_h = f->ws_w = 0;
}

struct FbGeom { int h1, w1, hp1, wp1, h2, w2, h3, w3, h4, w4, h5, w5; };
inline FbGeom fb_geom(int h, int w) {
  FbGeom g;
  g.h1 = conv_out(h, 7, 4, 3); g.w1 = conv_out(w, 7, 4, 3);
  g.hp1 = conv_out(g.h1, 3, 2, 1); g.wp1 = conv_out(g.w1, 3, 2, 1);
  g.h2 = conv_out(g.hp1, 5, 2, 2); g.w2 = conv_out(g.wp1, 5, 2, 2);
  g.h3 = conv_out(g.h2, 3, 2, 1); g.w3 = conv_out(g.w2, 3, 2, 1);
  g.h4 = conv_out(g.h3, 3, 2, 1); g.w4 = conv_out(g.w3, 3, 2, 1);
  g.h5 = conv_out(g.h4, 3, 2, 1); g.w5 = conv_out(g.w4, 3, 2, 1);
  return g;
}

int fb_workspace(syn_fb* f, int h, int w) {
  if (h == f->ws_h && w == f->ws_w) return SYN_OK;
  SYN_CUDA(cudaDeviceSynchronize());
  fb_free_ws(f);
  const FbGeom g = fb_geom(h, w);
  const size_t n3 = (size_t)g.h3 * g.w3, n4 = (size_t)g.h4 * g.w4, n5 = (size_t)g.h5 * g.w5;
  SYN_CUDA(cudaMalloc(&f->c1, sizeof(float) * g.h1 * g.w1 * 48));
  SYN_CUDA(cudaMalloc(&f->p1, sizeof(float) * g.hp1 * g.wp1 * 48));
  SYN_CUDA(cudaMalloc(&f->c2, sizeof(float) * g.h2 * g.w2 * 128));
  SYN_CUDA(cudaMalloc(&f->xa, sizeof(float) * n3 * 128));
  SYN_CUDA(cudaMalloc(&f->xb, sizeof(float) * n3 * 128));
  SYN_CUDA(cudaMalloc(&f->avg, sizeof(float) * n3 * 128));
  SYN_CUDA(cudaMalloc(&f->r1, sizeof(float) * n3 * 24));
  SYN_CUDA(cudaMalloc(&f->r2, sizeof(float) * n3 * 24));
  SYN_CUDA(cudaMalloc(&f->t3, sizeof(float) * n3 * 32));
  SYN_CUDA(cudaMalloc(&f->c31, sizeof(float) * n3 * 128));
  SYN_CUDA(cudaMalloc(&f->c32, sizeof(float) * n4 * 256));
  SYN_CUDA(cudaMalloc(&f->c41, sizeof(float) * n4 * 128));
  SYN_CUDA(cudaMalloc(&f->c42, sizeof(float) * n5 * 256));
  f->ws_h = h; f->ws_w = w;
  return SYN_OK;
}

int fb_conv(syn_fb* f, int idx, const float* x, const uint8_t* x_u8, int h, int w, int cin_stride, int cin_off, float* y,
            int cout_stride, int cout_off, cudaStream_t st) {
  const FbLayer& L = kFbLayers[idx];
  FbConvArgs a;
  a.x = x; a.x_u8 = x_u8; a.wk = f->d_w[idx]; a.bias = f->d_b[idx]; a.y = y;
  a.h = h; a.w = w; a.cin = L.cin; a.cin_stride = cin_stride; a.cin_off = cin_off;
  a.ho = conv_out(h, L.k, L.stride, L.pad); a.wo = conv_out(w, L.k, L.stride, L.pad);
  a.cout = L.cout; a.cout_stride = cout_stride; a.cout_off = cout_off;
  a.k = L.k; a.stride = L.stride; a.pad = L.pad; a.act = L.act;
  a.mean[0] = 104.f; a.mean[1] = 117.f; a.mean[2] = 123.f;          // FaceBoxes.py:92
  const int M = a.ho * a.wo;
  const dim3 grid((M + FB_BM - 1) / FB_BM, (L.cout + FB_BN - 1) / FB_BN);
  const bool vec = x_u8 == nullptr && L.cin % 4 == 0 && cin_stride % 4 == 0 && cin_off % 4 == 0 && L.cout % 4 == 0 &&
                   (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  if (L.cout <= FB_SMALLN && L.k * L.k * L.cin >= 512) fb_conv_smalln_kernel<<<M, 128, 0, st>>>(a);
  else if (vec) fb_conv_kernel<true><<<grid, 256, 0, st>>>(a);
  else fb_conv_kernel<false><<<grid, 256, 0, st>>>(a);
  SYN_LAUNCH_CHECK("fb_conv_kernel");
  ++f->launches;
  return SYN_OK;
}

}  // namespace

extern "C" {

int syn_fb_num_layers(void) { return 33; }

int syn_fb_layer_desc(int idx, syn_fb_layer_desc_t* out) {
  if (idx < 0 || idx >= 33 || !out) return fail(SYN_ERR_INVALID, "syn_fb_layer_desc: bad index %d", idx);
  const FbLayer& L = kFbLayers[idx];
  out->name = L.name; out->cin = L.cin; out->cout = L.cout; out->ksize = L.k; out->stride = L.stride; out->pad = L.pad;
  out->has_bn = L.bn; out->activation = L.act;
  return SYN_OK;
}

int syn_fb_create(int device, syn_fb_t** out) {
  if (!out) return fail(SYN_ERR_INVALID, "syn_fb_create: null out");
  SYN_CUDA(cudaSetDevice(device));
  syn_fb* f = new (std::nothrow) syn_fb();
  if (!f) return fail(SYN_ERR_NOMEM, "syn_fb_create: out of host memory");
  f->device = device;
  *out = f;
  return SYN_OK;
}

void syn_fb_destroy(syn_fb_t* f) {
  if (!f) return;
  cudaSetDevice(f->device);
  fb_free_ws(f);
  for (int i = 0; i < 33; ++i) { cudaFree(f->d_w[i]); cudaFree(f->d_b[i]); }
  delete f;
}

int syn_fb_set_layer(syn_fb_t* f, int idx, const float* w_host, int64_t w_numel, const float* bias_host, const float* bn_weight_host,
                     const float* bn_bias_host, const float* bn_mean_host, const float* bn_var_host, float eps) {
  if (!f || !w_host || idx < 0 || idx >= 33) return fail(SYN_ERR_INVALID, "syn_fb_set_layer: bad argument");
  const FbLayer& L = kFbLayers[idx];
  const int64_t want = (int64_t)L.cout * L.cin * L.k * L.k;
  if (w_numel != want) return fail(SYN_ERR_SHAPE, "syn_fb_set_layer: %s expects %lld weights, got %lld", L.name, (long long)want, (long long)w_numel);
  if (L.bn && (!bn_weight_host || !bn_bias_host || !bn_mean_host || !bn_var_host)) return fail(SYN_ERR_INVALID, "syn_fb_set_layer: %s needs its BatchNorm", L.name);
  if (!L.bn && !bias_host) return fail(SYN_ERR_INVALID, "syn_fb_set_layer: %s needs its bias", L.name);
  const int K = L.k * L.k * L.cin;
  f->w[idx].assign((size_t)K * L.cout, 0.f);
  f->b[idx].assign(L.cout, 0.f);
  for (int co = 0; co < L.cout; ++co) {
    double sc = 1.0, sh = 0.0;
    if (L.bn) {                                            // eval BatchNorm2d: y = (x - mean) / sqrt(var + eps) * weight + bias
      sc = (double)bn_weight_host[co] / std::sqrt((double)bn_var_host[co] + (double)eps);
      sh = (double)bn_bias_host[co] - (double)bn_mean_host[co] * sc;
    } else {
      sh = bias_host[co];
    }
    f->b[idx][co] = (float)sh;
    for (int ci = 0; ci < L.cin; ++ci)
      for (int kh = 0; kh < L.k; ++kh)
        for (int kw = 0; kw < L.k; ++kw)                   // OIHW -> [(kh, kw, ci)][co]
          f->w[idx][((size_t)(kh * L.k + kw) * L.cin + ci) * L.cout + co] =
              (float)((double)w_host[(((size_t)co * L.cin + ci) * L.k + kh) * L.k + kw] * sc);
  }
  f->set[idx] = true;
  f->committed = false;
  return SYN_OK;
}

int syn_fb_commit(syn_fb_t* f) {
  if (!f) return fail(SYN_ERR_INVALID, "syn_fb_commit: null handle");
  SYN_CUDA(cudaSetDevice(f->device));
  for (int i = 0; i < 33; ++i)
    if (!f->set[i]) return fail(SYN_ERR_STATE, "syn_fb_commit: layer %s was never set", kFbLayers[i].name);
  for (int i = 0; i < 33; ++i) {
    cudaFree(f->d_w[i]); cudaFree(f->d_b[i]);
    f->d_w[i] = f->d_b[i] = nullptr;
    SYN_CUDA(cudaMalloc(&f->d_w[i], f->w[i].size() * sizeof(float)));
    SYN_CUDA(cudaMalloc(&f->d_b[i], f->b[i].size() * sizeof(float)));
    SYN_CUDA(cudaMemcpy(f->d_w[i], f->w[i].data(), f->w[i].size() * sizeof(float), cudaMemcpyHostToDevice));
    SYN_CUDA(cudaMemcpy(f->d_b[i], f->b[i].data(), f->b[i].size() * sizeof(float), cudaMemcpyHostToDevice));
  }
  f->committed = true;
  return SYN_OK;
}

int64_t syn_fb_launch_count(const syn_fb_t* f) { return f ? f->launches : 0; }

int syn_fb_forward(syn_fb_t* f, const uint8_t* image_dev, int height, int width, float* loc_dev, float* conf_dev, void* stream) {
  if (!f || !image_dev || !loc_dev || !conf_dev || height <= 0 || width <= 0) return fail(SYN_ERR_INVALID, "syn_fb_forward: bad argument");
  if (!f->committed) return fail(SYN_ERR_STATE, "syn_fb_forward before syn_fb_commit");
  SYN_CUDA(cudaSetDevice(f->device));
  if (int rc = fb_workspace(f, height, width)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const FbGeom g = fb_geom(height, width);
  if (g.h3 != fb_cells(height, 32) || g.w3 != fb_cells(width, 32) || g.h4 != fb_cells(height, 64) || g.w4 != fb_cells(width, 64) ||
      g.h5 != fb_cells(height, 128) || g.w5 != fb_cells(width, 128))
    return fail(SYN_ERR_SHAPE, "syn_fb_forward: feature maps of a %dx%d input do not match the prior grid", height, width);
  auto pool_grid = [](size_t n) { return (unsigned)((n + 255) / 256); };
  // conv1 (CReLU) -> max-pool -> conv2 (CReLU) -> max-pool                                          faceboxes.py:120-123
  if (int rc = fb_conv(f, 0, nullptr, image_dev, height, width, 3, 0, f->c1, 48, 0, st)) return rc;
  fb_maxpool_kernel<<<pool_grid((size_t)g.hp1 * g.wp1 * 48), 256, 0, st>>>(f->c1, g.h1, g.w1, 48, f->p1, g.hp1, g.wp1);
  SYN_LAUNCH_CHECK("fb_maxpool_kernel");
  if (int rc = fb_conv(f, 1, f->p1, nullptr, g.hp1, g.wp1, 48, 0, f->c2, 128, 0, st)) return rc;
  fb_maxpool_kernel<<<pool_grid((size_t)g.h3 * g.w3 * 128), 256, 0, st>>>(f->c2, g.h2, g.w2, 128, f->xa, g.h3, g.w3);
  SYN_LAUNCH_CHECK("fb_maxpool_kernel");
  f->launches += 2;
  // three inception blocks: every branch writes its 32-channel slice of the next 128-channel tensor     :124-126, :33-47
  float *x = f->xa, *y = f->xb;
  for (int blk = 0; blk < 3; ++blk) {
    const int L0 = 2 + 7 * blk;
    if (int rc = fb_conv(f, L0 + 0, x, nullptr, g.h3, g.w3, 128, 0, y, 128, 0, st)) return rc;
    fb_avgpool_kernel<<<pool_grid((size_t)g.h3 * g.w3 * 128), 256, 0, st>>>(x, g.h3, g.w3, 128, f->avg);
    SYN_LAUNCH_CHECK("fb_avgpool_kernel");
    ++f->launches;
    if (int rc = fb_conv(f, L0 + 1, f->avg, nullptr, g.h3, g.w3, 128, 0, y, 128, 32, st)) return rc;
    if (int rc = fb_conv(f, L0 + 2, x, nullptr, g.h3, g.w3, 128, 0, f->r1, 24, 0, st)) return rc;
    if (int rc = fb_conv(f, L0 + 3, f->r1, nullptr, g.h3, g.w3, 24, 0, y, 128, 64, st)) return rc;
    if (int rc = fb_conv(f, L0 + 4, x, nullptr, g.h3, g.w3, 128, 0, f->r2, 24, 0, st)) return rc;
    if (int rc = fb_conv(f, L0 + 5, f->r2, nullptr, g.h3, g.w3, 24, 0, f->t3, 32, 0, st)) return rc;
    if (int rc = fb_conv(f, L0 + 6, f->t3, nullptr, g.h3, g.w3, 32, 0, y, 128, 96, st)) return rc;
    float* t = x; x = y; y = t;
  }
  // x = inception3 output (detection source 0); conv3_x, conv4_x give sources 1 and 2                  :127-135
  if (int rc = fb_conv(f, 23, x, nullptr, g.h3, g.w3, 128, 0, f->c31, 128, 0, st)) return rc;
  if (int rc = fb_conv(f, 24, f->c31, nullptr, g.h3, g.w3, 128, 0, f->c32, 256, 0, st)) return rc;
  if (int rc = fb_conv(f, 25, f->c32, nullptr, g.h4, g.w4, 256, 0, f->c41, 128, 0, st)) return rc;
  if (int rc = fb_conv(f, 26, f->c41, nullptr, g.h4, g.w4, 128, 0, f->c42, 256, 0, st)) return rc;
  // heads: NHWC output of each source IS permute(0,2,3,1).view(-1) (:137-142); the three sources are concatenated by offset
  const size_t n3 = (size_t)g.h3 * g.w3, n4 = (size_t)g.h4 * g.w4, n5 = (size_t)g.h5 * g.w5;
  if (int rc = fb_conv(f, 27, x, nullptr, g.h3, g.w3, 128, 0, loc_dev, 84, 0, st)) return rc;
  if (int rc = fb_conv(f, 28, f->c32, nullptr, g.h4, g.w4, 256, 0, loc_dev + n3 * 84, 4, 0, st)) return rc;
  if (int rc = fb_conv(f, 29, f->c42, nullptr, g.h5, g.w5, 256, 0, loc_dev + n3 * 84 + n4 * 4, 4, 0, st)) return rc;
  if (int rc = fb_conv(f, 30, x, nullptr, g.h3, g.w3, 128, 0, conf_dev, 42, 0, st)) return rc;
  if (int rc = fb_conv(f, 31, f->c32, nullptr, g.h4, g.w4, 256, 0, conf_dev + n3 * 42, 2, 0, st)) return rc;
  if (int rc = fb_conv(f, 32, f->c42, nullptr, g.h5, g.w5, 256, 0, conf_dev + n3 * 42 + n4 * 2, 2, 0, st)) return rc;
  const int np = (int)(n3 * 21 + n4 + n5);
  fb_softmax2_kernel<<<(np + 255) / 256, 256, 0, st>>>(conf_dev, np);
  SYN_LAUNCH_CHECK("fb_softmax2_kernel");
  ++f->launches;
  return SYN_OK;
}

}  // extern "C"
