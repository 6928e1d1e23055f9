/*
 * C restatement of the byte/index-exact pieces of the scene pipeline -- TEST INFRASTRUCTURE.
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load this; the
 * product never does.  Each function cites the reference lines it follows (paths relative to
 * /root/reference/src).  Checked against the golden vectors in tests/test_oracle.py.
 *
 * Build: make -C oracle   (-> oracle/_build/liboracle.so)
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

static const int VIEW_ORDER[6] = {0, 1, 2, 5, 4, 3}; /* roadmap_bce_v2.py:58 */

/* roadmap_bce_v2.py:53-64: mosaic[b,c,h,j*W+w] = views[b,VIEW_ORDER[j],c,h,w] */
void oracle_stitch(const float* views, float* mosaic, int B, int H, int W) {
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < 3; ++c)
      for (int h = 0; h < H; ++h)
        for (int j = 0; j < 6; ++j)
          memcpy(mosaic + (((size_t)b * 3 + c) * H + h) * 6 * W + (size_t)j * W,
                 views + ((((size_t)b * 6 + VIEW_ORDER[j]) * 3 + c) * H + h) * W, sizeof(float) * W);
}

/* autoencoder.py:53-73 with the drawn slot passed in */
void oracle_stitch_mask(const float* views, float* x, float* y, int B, int H, int W, int slot) {
  oracle_stitch(views, x, B, H, W);
  for (size_t row = 0; row < (size_t)B * 3 * H; ++row) {
    float* xr = x + row * 6 * W + (size_t)slot * W;
    memcpy(y + row * W, xr, sizeof(float) * W);
    memset(xr, 0, sizeof(float) * W);
  }
}

/* components.py:46-47: max over 4 consecutive NCHW-flat elements; argmax = first maximum */
void oracle_pool4_flat(const float* a3_nchw, float* pooled, uint8_t* argmax, int B, long long n) {
  const long long q = n / 4;
  for (int b = 0; b < B; ++b)
    for (long long j = 0; j < q; ++j) {
      const float* p = a3_nchw + (size_t)b * n + 4 * j;
      int m = 0;
      for (int e = 1; e < 4; ++e)
        if (p[e] > p[m]) m = e;
      pooled[(size_t)b * q + j] = p[m];
      if (argmax) argmax[(size_t)b * q + j] = (uint8_t)m;
    }
}

/* components.py:19-21,41-43: 3x3 conv, pad 1, stride s, + bias + ReLU, NCHW, double accumulate */
void oracle_conv3x3_relu(const float* x, const float* w, const float* bias, float* y, int B, int Cin, int Cout,
                         int H, int W, int stride) {
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  for (int b = 0; b < B; ++b)
    for (int co = 0; co < Cout; ++co)
      for (int ho = 0; ho < Ho; ++ho)
        for (int wo = 0; wo < Wo; ++wo) {
          double acc = bias[co];
          for (int ci = 0; ci < Cin; ++ci)
            for (int kh = 0; kh < 3; ++kh)
              for (int kw = 0; kw < 3; ++kw) {
                const int h = ho * stride + kh - 1, ww = wo * stride + kw - 1;
                if (h < 0 || h >= H || ww < 0 || ww >= W) continue;
                acc += (double)x[(((size_t)b * Cin + ci) * H + h) * W + ww] *
                       (double)w[(((size_t)co * Cin + ci) * 3 + kh) * 3 + kw];
              }
          y[(((size_t)b * Cout + co) * Ho + ho) * Wo + wo] = acc > 0.0 ? (float)acc : 0.0f;
        }
}

/* roadmap_bce_v2.py:140: probs.round() from logits == (x > 1.5*2^-24), see scene_oracle.binarise */
void oracle_binarise(const float* logits, uint8_t* out, long long n) {
  union { uint32_t u; float f; } thr;
  thr.u = 0x33C00000u;
  for (long long i = 0; i < n; ++i) out[i] = logits[i] > thr.f;
}

/* helper.py:74-77 on 0/1 maps, as exact integer counts: {sum t, sum r, sum t*r} */
void oracle_ts_counts(const uint8_t* target, const uint8_t* pred, long long n, long long* out3) {
  long long nt = 0, nr = 0, ntr = 0;
  for (long long i = 0; i < n; ++i) {
    nt += target[i] != 0;
    nr += pred[i] != 0;
    ntr += (target[i] != 0) & (pred[i] != 0);
  }
  out3[0] = nt; out3[1] = nr; out3[2] = ntr;
}

/* roadmap_bce_v2.py:103-106: mean of (1-t)x + max(-x,0) + log1p(exp(-|x|)), double accumulate */
double oracle_bce_mean(const float* logits, const float* target, long long n) {
  double s = 0.0;
  for (long long i = 0; i < n; ++i) {
    const double x = logits[i], t = target[i];
    s += (1.0 - t) * x + (x < 0 ? -x : 0.0) + log1p(exp(-fabs(x)));
  }
  return s / (double)n;
}
