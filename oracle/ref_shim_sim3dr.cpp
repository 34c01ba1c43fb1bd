// extern "C" doorway into the reference's own Sim3DR translation unit, compiled where it lies under /root/reference
// (oracle/Makefile; output oracle/_ref/libsim3dr_ref.so, git-ignored).  No reference source is copied: this file only
// forwards to the functions rasterize.h declares.
#include "rasterize.h"

extern "C" {
void ref_get_normal(float* ver_normal, float* vertices, int* triangles, int nver, int ntri) {
  _get_normal(ver_normal, vertices, triangles, nver, ntri);
}
void ref_rasterize(unsigned char* image, float* vertices, int* triangles, float* colors, float* depth_buffer, int ntri, int h,
                   int w, int c, float alpha, int reverse) {
  _rasterize(image, vertices, triangles, colors, depth_buffer, ntri, h, w, c, alpha, reverse != 0);
}
}
