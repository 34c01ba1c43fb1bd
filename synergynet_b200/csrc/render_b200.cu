// C ABI of the stages either side of the 3DMM path: Sim3DR normals / lighting /
// rasterisation of the dense meshes, and the FaceBoxes box decode + greedy NMS that produces the crops.
// Handle-free: device pointers and workspaces belong to the caller (include/synergy_b200.h states the sizes).
#include "kernels_render.cuh"
#include "kernels_detect.cuh"

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
  nms_scan_kernel<<<1, 32, words * 8, st>>>(reinterpret_cast<const unsigned long long*>(mask_ws_dev), n, keep_dev, n_keep_dev);
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
