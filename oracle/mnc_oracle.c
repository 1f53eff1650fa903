/*
 * oracle/mnc_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement (OpenMP over independent RoIs where the loop is embarrassingly parallel) of the CUDA kernels on the MNC inference hot path
 * of the reference (daijifeng001/MNC).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; mnc_b200/ never does.
 *
 * Each function cites the reference file:line it follows and mirrors its expressions term by
 * term in C (same int/float/double promotions), compiled with -ffp-contract=off so no FMA
 * contraction happens behind our back.
 *
 * Pinning: the reference ships no forward known-answer vectors for these kernels
 * (SURVEY.md section 8c), so every function here is pinned against the reference's own SOURCES
 * compiled unmodified into oracle/_ref (oracle/Makefile `ref`; tests/test_ref_pin.py on the GPU):
 * orc_nms / orc_mv against lib/nms/{nms,mv}_kernel.cu; orc_roi_warp / orc_mask_resize /
 * orc_mask_pool / orc_roi_pool against caffe-mnc/src/caffe/layers/{roi_warping,mask_resize,
 * mask_pooling,roi_pooling}_layer.{cu,cpp} (built against the Caffe-runtime stand-in
 * oracle/ref_stub).  Built with -fmad=false the reference equals this file BIT FOR BIT; built with
 * nvcc's default FMA contraction it differs by the rounding of one fused product (<= 2e-5).
 * orc_roi_pool is additionally checked on the CPU against the reference's ROIPooling Forward_cpu
 * (tests/test_ref_fixtures.py).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------------------------------
 * IoU with the +1 pixel convention -- lib/nms/nms_kernel.cu:24-32 (devIoU). */
static float dev_iou(const float* a, const float* b) {
  float left = fmaxf(a[0], b[0]), right = fminf(a[2], b[2]);
  float top = fmaxf(a[1], b[1]), bottom = fminf(a[3], b[3]);
  float width = fmaxf(right - left + 1, 0.f), height = fmaxf(bottom - top + 1, 0.f);
  float interS = width * height;
  float Sa = (a[2] - a[0] + 1) * (a[3] - a[1] + 1);
  float Sb = (b[2] - b[0] + 1) * (b[3] - b[1] + 1);
  return interS / (Sa + Sb - interS);
}

/* Greedy NMS on score-sorted boxes, suppress when IoU > thresh (strict) --
 * lib/nms/nms_kernel.cu:61-77 (mask rows: j > i, IoU > thr) + :124-139 (host greedy reduce).
 * boxes: n x box_dim floats (x1,y1,x2,y2,[score]); keep_out: >= n ints. */
void orc_nms(const float* boxes, int n, int box_dim, float thresh, int* keep_out, int* num_out) {
  unsigned char* removed = (unsigned char*)calloc((size_t)(n > 0 ? n : 1), 1);
  int num = 0;
  for (int i = 0; i < n; ++i) {
    if (removed[i]) continue;
    keep_out[num++] = i;
    const float* bi = boxes + (size_t)i * box_dim;
    for (int j = i + 1; j < n; ++j) {
      if (removed[j]) continue; /* OR-ing an already-set bit changes nothing */
      if (dev_iou(bi, boxes + (size_t)j * box_dim) > thresh) removed[j] = 1;
    }
  }
  *num_out = num;
  free(removed);
}

/* float64 IoU matrix -- lib/utils/bbox.pyx:15-55 (bbox_overlaps). out: N x K row-major. */
void orc_bbox_overlaps(const double* boxes, int N, const double* query, int K, double* out) {
  memset(out, 0, sizeof(double) * (size_t)N * K);
  for (int k = 0; k < K; ++k) {
    const double* q = query + 4 * k;
    double box_area = (q[2] - q[0] + 1) * (q[3] - q[1] + 1);
    for (int n = 0; n < N; ++n) {
      const double* b = boxes + 4 * n;
      double iw = fmin(b[2], q[2]) - fmax(b[0], q[0]) + 1;
      if (iw > 0) {
        double ih = fmin(b[3], q[3]) - fmax(b[1], q[1]) + 1;
        if (ih > 0) {
          double ua = (b[2] - b[0] + 1) * (b[3] - b[1] + 1) + box_area - iw * ih;
          out[(size_t)n * K + k] = iw * ih / ua;
        }
      }
    }
  }
}

/* ---------------------------------------------------------------------------------------
 * Bilinear sample -- caffe-mnc/src/caffe/layers/roi_warping_layer.cu:18-64.  Returns 0 and
 * *hit = 0 when the sample is outside [-0.5, dim-0.5] (the reference leaves maxidx = -1 and the
 * caller writes 0, :102). */
static float warp_bilinear(const float* data, int height, int width, float h, float w, int* hit) {
  if (h < -0.5 || h > height - 0.5 || w < -0.5 || w > width - 0.5) {
    *hit = 0;
    return 0.f;
  }
  *hit = 1;
  if (h <= 0) h = 0;
  if (w <= 0) w = 0;
  int h_low = (int)h, w_low = (int)w, h_high, w_high;
  if (h_low >= height - 1) {
    h_high = h_low = height - 1;
    h = (float)h_low;
  } else {
    h_high = h_low + 1;
  }
  if (w_low >= width - 1) {
    w_high = w_low = width - 1;
    w = (float)w_low;
  } else {
    w_high = w_low + 1;
  }
  float lh = h - h_low, lw = w - w_low;
  float hh = 1 - lh, hw = 1 - lw;
  float v1 = data[h_low * width + w_low], v2 = data[h_low * width + w_high];
  float v3 = data[h_high * width + w_low], v4 = data[h_high * width + w_high];
  float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
  return (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4);
}

/* ROIWarping forward -- roi_warping_layer.cu:67-107.  feat (B,C,H,W) NCHW, rois (R,5)
 * [batch_idx,x1,y1,x2,y2], out (R,C,ph,pw). */
void orc_roi_warp(const float* feat, int C, int H, int W, const float* rois, int R, int ph_n,
                  int pw_n, float spatial_scale, float* out) {
#pragma omp parallel for schedule(dynamic, 4) /* RoIs are independent; per-element math unchanged */
  for (int n = 0; n < R; ++n) {
    const float* roi = rois + 5 * n;
    int roi_level = (int)roi[0];
    float roi_start_w = roundf(roi[1] * spatial_scale);
    float roi_start_h = roundf(roi[2] * spatial_scale);
    float roi_end_w = roundf(roi[3] * spatial_scale);
    float roi_end_h = roundf(roi[4] * spatial_scale);
    float roi_width = fmaxf(roi_end_w - roi_start_w, 0.f);
    float roi_height = fmaxf(roi_end_h - roi_start_h, 0.f);
    float bin_size_h = roi_height / (float)ph_n;
    float bin_size_w = roi_width / (float)pw_n;
    for (int c = 0; c < C; ++c) {
      const float* plane = feat + ((size_t)roi_level * C + c) * H * W;
      for (int ph = 0; ph < ph_n; ++ph)
        for (int pw = 0; pw < pw_n; ++pw) {
          float ih = roi_start_h + (float)ph * bin_size_h;
          float iw = roi_start_w + (float)pw * bin_size_w;
          int hit;
          float v = warp_bilinear(plane, H, W, ih, iw, &hit);
          /* maxval starts at -FLT_MAX; `val > maxval` (:58) holds for every finite val */
          out[(((size_t)n * C + c) * ph_n + ph) * pw_n + pw] = hit ? v : 0.f;
        }
    }
  }
}

/* ROIPooling forward -- caffe-mnc/src/caffe/layers/roi_pooling_layer.cu:17-77 (same arithmetic as
 * the CPU form roi_pooling_layer.cpp:47-127).  feat (B,C,H,W), rois (R,5), out/argmax (R,C,ph,pw);
 * argmax may be NULL. */
void orc_roi_pool(const float* feat, int C, int H, int W, const float* rois, int R, int ph_n,
                  int pw_n, float spatial_scale, float* out, int* argmax) {
#pragma omp parallel for schedule(dynamic, 4)
  for (int n = 0; n < R; ++n) {
    const float* roi = rois + 5 * n;
    int roi_batch_ind = (int)roi[0];
    int roi_start_w = (int)roundf(roi[1] * spatial_scale);
    int roi_start_h = (int)roundf(roi[2] * spatial_scale);
    int roi_end_w = (int)roundf(roi[3] * spatial_scale);
    int roi_end_h = (int)roundf(roi[4] * spatial_scale);
    int roi_width = roi_end_w - roi_start_w + 1 > 1 ? roi_end_w - roi_start_w + 1 : 1;
    int roi_height = roi_end_h - roi_start_h + 1 > 1 ? roi_end_h - roi_start_h + 1 : 1;
    float bin_size_h = (float)roi_height / (float)ph_n;
    float bin_size_w = (float)roi_width / (float)pw_n;
    for (int c = 0; c < C; ++c) {
      const float* plane = feat + ((size_t)roi_batch_ind * C + c) * H * W;
      for (int ph = 0; ph < ph_n; ++ph)
        for (int pw = 0; pw < pw_n; ++pw) {
          int hstart = (int)floorf((float)ph * bin_size_h);
          int wstart = (int)floorf((float)pw * bin_size_w);
          int hend = (int)ceilf((float)(ph + 1) * bin_size_h);
          int wend = (int)ceilf((float)(pw + 1) * bin_size_w);
          hstart = hstart + roi_start_h < 0 ? 0 : (hstart + roi_start_h > H ? H : hstart + roi_start_h);
          hend = hend + roi_start_h < 0 ? 0 : (hend + roi_start_h > H ? H : hend + roi_start_h);
          wstart = wstart + roi_start_w < 0 ? 0 : (wstart + roi_start_w > W ? W : wstart + roi_start_w);
          wend = wend + roi_start_w < 0 ? 0 : (wend + roi_start_w > W ? W : wend + roi_start_w);
          int is_empty = (hend <= hstart) || (wend <= wstart);
          float maxval = is_empty ? 0.f : -3.402823466e+38f;
          int maxidx = -1;
          for (int h = hstart; h < hend; ++h)
            for (int w = wstart; w < wend; ++w)
              if (plane[h * W + w] > maxval) {
                maxval = plane[h * W + w];
                maxidx = h * W + w;
              }
          size_t o = (((size_t)n * C + c) * ph_n + ph) * pw_n + pw;
          out[o] = maxval;
          if (argmax) argmax[o] = maxidx;
        }
    }
  }
}

/* MaskResize forward -- caffe-mnc/src/caffe/layers/mask_resize_layer.cu:13-73. */
void orc_mask_resize(const float* in, int N, int C, int ih_n, int iw_n, int oh_n, int ow_n,
                     float* out) {
  float ratio_h = (float)ih_n / (float)oh_n;
  float ratio_w = (float)iw_n / (float)ow_n;
#pragma omp parallel for
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < C; ++c) {
      const float* plane = in + ((size_t)n * C + c) * ih_n * iw_n;
      for (int h = 0; h < oh_n; ++h)
        for (int w = 0; w < ow_n; ++w) {
          float inverse_x = w * ratio_w;
          float inverse_y = h * ratio_h;
          int hit;
          float v = warp_bilinear(plane, ih_n, iw_n, inverse_y, inverse_x, &hit);
          out[(((size_t)n * C + c) * oh_n + h) * ow_n + w] = hit ? v : 0.f;
        }
    }
}

/* MaskPooling forward -- caffe-mnc/src/caffe/layers/mask_pooling_layer.cu:13-26. */
void orc_mask_pool(const float* feat, const float* mask, int N, int C, int H, int W, float* out) {
#pragma omp parallel for
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < C; ++c)
      for (int i = 0; i < H * W; ++i)
        out[((size_t)n * C + c) * H * W + i] =
            feat[((size_t)n * C + c) * H * W + i] * mask[(size_t)n * H * W + i];
}

/* ---------------------------------------------------------------------------------------
 * Mask voting device part -- lib/nms/mv_kernel.cu. */
static const float BINARIZE_THRESH = 0.4; /* mv_kernel.cu:13 */

/* mask_render (mv_kernel.cu:36-91) for one pixel of one box. */
static float mv_render(const float* box, const float* mask, int mask_size, int h, int w) {
  const float box_x1 = box[0], box_y1 = box[1], box_x2 = box[2], box_y2 = box[3];
  if (w < box_x1 || w > box_x2 || h < box_y1 || h > box_y2) return 0.0;
  const float box_width = box_x2 - box_x1 + 1.0;
  const float box_height = box_y2 - box_y1 + 1.0;
  const float ratio_w = (float)mask_size / box_width;
  const float ratio_h = (float)mask_size / box_height;
  const float inverse_x = ((float)w - box_x1) * ratio_w;
  const float inverse_y = ((float)h - box_y1) * ratio_h;
  int start_x = floor(inverse_x);
  int start_y = floor(inverse_y);
  if (start_x == mask_size - 1 && start_y == mask_size - 1) {
    return mask[mask_size * mask_size - 1];
  } else if (start_x == mask_size - 1 && start_y != mask_size - 1) {
    return mask[start_y * mask_size + start_x];
  } else if (start_x != mask_size - 1 && start_y == mask_size - 1) {
    return mask[start_y * mask_size + start_x];
  } else {
    int tl = start_y * mask_size + start_x, tr = tl + 1, bl = tl + mask_size, br = bl + 1;
    float tlw = (1 - (inverse_x - start_x)) * (1 - (inverse_y - start_y));
    float trw = (inverse_x - start_x) * (1 - (inverse_y - start_y));
    float blw = (1 - (inverse_x - start_x)) * (inverse_y - start_y);
    float brw = (inverse_x - start_x) * (inverse_y - start_y);
    float val = tlw * mask[tl] + trw * mask[tr] + blw * mask[bl] + brw * mask[br];
    return val;
  }
}

/* _mv (mv_kernel.cu:242-348) without the nb*H*W render buffer: one aggregated H*W image per
 * result at a time.  candidate_start holds END offsets (mv_kernel.cu:101-102).
 * If agg_out != NULL it receives the result_num aggregated images (result_num*H*W floats). */
void orc_mv(const float* all_boxes, const float* all_masks, int all_boxes_num,
            const int* candidate_inds, const int* candidate_start, const float* candidate_weights,
            int candidate_num, int image_height, int image_width, int box_dim, int mask_size,
            int result_num, float* out_mask, int* out_box, float* agg_out) {
  (void)all_boxes_num;
  (void)candidate_num;
  const int H = image_height, W = image_width, M = mask_size;
  float* agg = (float*)malloc(sizeof(float) * (size_t)H * W);
  for (int n = 0; n < result_num; ++n) {
    /* mask_aggregate :93-112 -- per pixel, sum in list order; boxes not covering a pixel add 0 */
    for (size_t i = 0; i < (size_t)H * W; ++i) agg[i] = 0.0;
    int cs = (n == 0) ? 0 : candidate_start[n - 1];
    int ce = candidate_start[n];
    for (int i = cs; i < ce; ++i) {
      int ind = candidate_inds[i];
      const float* box = all_boxes + (size_t)ind * box_dim;
      const float* msk = all_masks + (size_t)ind * M * M;
      float wgt = candidate_weights[i];
      int h_lo = (int)floorf(box[1]), h_hi = (int)ceilf(box[3]);
      int w_lo = (int)floorf(box[0]), w_hi = (int)ceilf(box[2]);
      if (h_lo < 0) h_lo = 0;
      if (w_lo < 0) w_lo = 0;
      if (h_hi > H - 1) h_hi = H - 1;
      if (w_hi > W - 1) w_hi = W - 1;
      for (int h = h_lo; h <= h_hi; ++h)
        for (int w = w_lo; w <= w_hi; ++w) {
          float r = mv_render(box, msk, M, h, w);
          agg[(size_t)h * W + w] += (r * wgt);
        }
    }
    if (agg_out) memcpy(agg_out + (size_t)n * H * W, agg, sizeof(float) * (size_t)H * W);
    /* reduce_mask_col/row :114-142 and reduce_bounding_x/y :144-190 */
    int bx0 = W / 2, bx1 = W / 2, by0 = H / 2, by1 = H / 2;
    int found = 0;
    for (int w = 0; w < W && !found; ++w)
      for (int h = 0; h < H; ++h)
        if (agg[(size_t)h * W + w] > BINARIZE_THRESH) {
          bx0 = w;
          found = 1;
          break;
        }
    found = 0;
    for (int w = W - 1; w >= 0 && !found; --w)
      for (int h = 0; h < H; ++h)
        if (agg[(size_t)h * W + w] > BINARIZE_THRESH) {
          bx1 = w;
          found = 1;
          break;
        }
    found = 0;
    for (int h = 0; h < H && !found; ++h)
      for (int w = 0; w < W; ++w)
        if (agg[(size_t)h * W + w] > BINARIZE_THRESH) {
          by0 = h;
          found = 1;
          break;
        }
    found = 0;
    for (int h = H - 1; h >= 0 && !found; --h)
      for (int w = 0; w < W; ++w)
        if (agg[(size_t)h * W + w] > BINARIZE_THRESH) {
          by1 = h;
          found = 1;
          break;
        }
    out_box[n * 4 + 0] = bx0;
    out_box[n * 4 + 1] = by0;
    out_box[n * 4 + 2] = bx1;
    out_box[n * 4 + 3] = by1;
    /* mask_resize :193-240 */
    for (int h = 0; h < M; ++h)
      for (int w = 0; w < M; ++w) {
        int bbox_x1 = bx0, bbox_x2 = bx1, bbox_y1 = by0, bbox_y2 = by1;
        float bbox_width = bbox_x2 - bbox_x1 + 1.0;
        float bbox_height = bbox_y2 - bbox_y1 + 1.0;
        float ratio_w = bbox_width / (float)M;
        float ratio_h = bbox_height / (float)M;
        float inverse_x = bbox_x1 + (float)w * ratio_w;
        float inverse_y = bbox_y1 + (float)h * ratio_h;
        int start_x = floor(inverse_x);
        int start_y = floor(inverse_y);
        float v;
        if (start_x == W - 1 && start_y == H - 1) {
          v = agg[(size_t)W * H - 1];
        } else if (start_x == W - 1 && start_y != H - 1) {
          v = agg[(size_t)start_y * W + start_x];
        } else if (start_x != W - 1 && start_y == H - 1) {
          v = agg[(size_t)start_y * W + start_x];
        } else {
          size_t tl = (size_t)start_y * W + start_x, tr = tl + 1, bl = tl + W, br = bl + 1;
          float tlw = (1 - (inverse_x - start_x)) * (1 - (inverse_y - start_y));
          float trw = (inverse_x - start_x) * (1 - (inverse_y - start_y));
          float blw = (1 - (inverse_x - start_x)) * (inverse_y - start_y);
          float brw = (inverse_x - start_x) * (inverse_y - start_y);
          v = tlw * agg[tl] + trw * agg[tr] + blw * agg[bl] + brw * agg[br];
        }
        out_mask[((size_t)n * M + h) * M + w] = v;
      }
  }
  free(agg);
}
