/*
 * CPU oracle, plain C: scalar restatement of the multi-scale deformable attention forward as mmcv-full 1.3.17's
 * CUDA kernel computes it (ms_deformable_im2col_gpu_kernel + ms_deform_attn_im2col_bilinear; the op the reference
 * calls at projects/UniBEV/unibev_plugin/models/modules/spatial_cross_attention_img.py:432-435,
 * spatial_cross_attention_pts.py:439-442, decoder.py:324-327).  mmcv is not vendored under /root/reference, so this
 * follows the published algorithm:
 *   h_im = loc_y * H_l - 0.5, w_im = loc_x * W_l - 0.5;
 *   a sample contributes iff h_im > -1 && w_im > -1 && h_im < H_l && w_im < W_l;
 *   each of the four corners is bounds-checked on its own (zero padding);
 *   out[b, q, h*D + c] = sum_{l, p} attn[b, q, h, l, p] * bilinear(value[b, start_l + y*W_l + x, h, c]).
 *
 * TEST INFRASTRUCTURE ONLY: built by __graft_entry__.build() into oracle/_build/libmsda_core.so and loaded by
 * tests/ (and bench.py's CPU-baseline leg); nothing under unibev_b200/ links or calls it.
 * Accumulation is in double so the result is the reference value to within one fp32 rounding.
 */
#include <math.h>
#include <stdint.h>

void ub_oracle_msda_forward(const float* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                            const float* sampling_loc, const float* attn_weight, float* out, int B, int Nv, int H,
                            int D, int Nq, int L, int P) {
  for (int b = 0; b < B; ++b)
    for (int q = 0; q < Nq; ++q)
      for (int h = 0; h < H; ++h) {
        const int64_t item = ((int64_t)b * Nq + q) * H + h;
        for (int c = 0; c < D; ++c) {
          double acc = 0.0;
          for (int l = 0; l < L; ++l) {
            const int fh = (int)spatial_shapes[2 * l], fw = (int)spatial_shapes[2 * l + 1];
            const float* base = value + (((int64_t)b * Nv + level_start_index[l]) * H + h) * D + c;
            const int64_t pix = (int64_t)H * D; /* floats between neighbouring value tokens */
            for (int p = 0; p < P; ++p) {
              const int64_t s = (item * L + l) * P + p;
              const float x = sampling_loc[2 * s], y = sampling_loc[2 * s + 1], a = attn_weight[s];
              const float h_im = y * (float)fh - 0.5f, w_im = x * (float)fw - 0.5f;
              if (!(h_im > -1.f && w_im > -1.f && h_im < (float)fh && w_im < (float)fw)) continue;
              const int h_low = (int)floorf(h_im), w_low = (int)floorf(w_im);
              const int h_high = h_low + 1, w_high = w_low + 1;
              const float lh = h_im - (float)h_low, lw = w_im - (float)w_low, hh = 1.f - lh, hw = 1.f - lw;
              float v1 = 0.f, v2 = 0.f, v3 = 0.f, v4 = 0.f;
              if (h_low >= 0 && w_low >= 0) v1 = base[((int64_t)h_low * fw + w_low) * pix];
              if (h_low >= 0 && w_high <= fw - 1) v2 = base[((int64_t)h_low * fw + w_high) * pix];
              if (h_high <= fh - 1 && w_low >= 0) v3 = base[((int64_t)h_high * fw + w_low) * pix];
              if (h_high <= fh - 1 && w_high <= fw - 1) v4 = base[((int64_t)h_high * fw + w_high) * pix];
              const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
              acc += (double)a * ((double)w1 * v1 + (double)w2 * v2 + (double)w3 * v3 + (double)w4 * v4);
            }
          }
          out[item * D + c] = (float)acc;
        }
      }
}
