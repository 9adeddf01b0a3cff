/* oracle.c — plain-C restatement of the PASTA-GAN operator hot path.  TEST INFRASTRUCTURE ONLY (see oracle/ops_oracle.py).
 *
 * Scalar loops with double accumulation, written from the definitions the reference implements:
 *   orc_upfirdn2d   torch_utils/ops/upfirdn2d.py:169-208 (_upfirdn2d_ref) / upfirdn2d.cu:29-92 (the generic gather kernel)
 *   orc_bias_act    torch_utils/ops/bias_act.py:94-123 (_bias_act_ref) and the grad = 1 branch of bias_act.cu:54-142
 *   orc_conv2d      what conv2d_gradfix.conv2d computes (conv2d_gradfix.py:35-38): cross-correlation, or true convolution with flip_weight = 0
 *                   (conv2d_resample.py:35-36)
 * It is a second, independent statement of the algorithm next to the torch-CPU oracle: tests/test_oracle_c.py pins it to the same golden
 * vectors (the .npz files under tests/golden, generated from the unmodified reference).  Parity status: PINNED.  Nothing in pasta-gan_b200/ links or loads it.
 * Build: make -C oracle  ->  oracle/_build/liboracle.so
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

static int floordiv(int a, int b) { int q = a / b; if ((a % b != 0) && ((a < 0) != (b < 0))) q--; return q; }

/* y[n,c,oy,ox] = gain * sum_{i,j} k[i,j] * z[oy*downy + i, ox*downx + j],  z = pad(zero_insert(x)),  k = f flipped unless `flip`. */
int orc_upfirdn2d(const float* x, const float* f, float* y, int N, int C, int H, int W, int fh, int fw,
                  int upx, int upy, int downx, int downy, int padx0, int padx1, int pady0, int pady1, int flip, float gain) {
    const int outW = (W * upx + padx0 + padx1 - fw + downx) / downx;
    const int outH = (H * upy + pady0 + pady1 - fh + downy) / downy;
    if (outW < 1 || outH < 1) return 1;
#pragma omp parallel for collapse(2)
    for (int n = 0; n < N; n++)
        for (int c = 0; c < C; c++) {
            const float* xp = x + ((size_t)n * C + c) * H * W;
            float* yp = y + ((size_t)n * C + c) * outH * outW;
            for (int oy = 0; oy < outH; oy++)
                for (int ox = 0; ox < outW; ox++) {
                    double acc = 0.0;
                    for (int i = 0; i < fh; i++) {
                        const int uy = oy * downy + i - pady0;            /* row in the zero-inserted image */
                        if (uy < 0 || uy % upy != 0) continue;
                        const int iy = uy / upy;
                        if (iy >= H) continue;
                        for (int j = 0; j < fw; j++) {
                            const int ux = ox * downx + j - padx0;
                            if (ux < 0 || ux % upx != 0) continue;
                            const int ix = ux / upx;
                            if (ix >= W) continue;
                            const float kv = flip ? f[i * fw + j] : f[(fh - 1 - i) * fw + (fw - 1 - j)];
                            acc += (double)kv * (double)xp[(size_t)iy * W + ix];
                        }
                    }
                    yp[(size_t)oy * outW + ox] = (float)(acc * (double)gain);
                }
        }
    (void)floordiv;
    return 0;
}

static double act_fwd(int act, double t, double alpha) {
    switch (act) {
        case 1: return t;
        case 2: return t > 0 ? t : 0;
        case 3: return t > 0 ? t : t * alpha;
        case 4: return tanh(t);
        case 5: return 1.0 / (1.0 + exp(-t));
        case 6: return t >= 0 ? t : expm1(t);
        case 7: return 1.0507009873554804934193349852946 * (t >= 0 ? t : 1.6732632423543772848170429916717 * expm1(t));
        case 8: return t > 20 ? t : log1p(exp(t));
        case 9: return t / (1.0 + exp(-t));
    }
    return NAN;
}

/* grad == 0: y = clamp(act(x + b) * gain).   grad == 1: y = x * gain * act'(.) from yref (linear / relu / lrelu only), 0 where |yref| >= clamp. */
int orc_bias_act(const float* x, const float* b, const float* yref, float* y, int64_t n, int size_b, int64_t step_b,
                 int grad, int act, float alpha, float gain, float clamp) {
    if (grad == 1 && act > 3) return 2;
#pragma omp parallel for
    for (int64_t i = 0; i < n; i++) {
        const double bias = b ? (double)b[(i / step_b) % size_b] : 0.0;
        double v;
        if (grad == 0) {
            v = act_fwd(act, (double)x[i] + bias, alpha) * gain;
            if (clamp >= 0) v = v > clamp ? clamp : (v < -clamp ? -clamp : v);
        } else {
            const double yy = gain != 0 ? (double)yref[i] / gain : 0.0;
            const double d1 = act == 1 ? 1.0 : (act == 2 ? (yy > 0 ? 1.0 : 0.0) : (yy > 0 ? 1.0 : alpha));
            v = (double)x[i] * d1 * gain;
            if (clamp >= 0 && !((double)yref[i] > -clamp && (double)yref[i] < clamp)) v = 0.0;
        }
        y[i] = (float)v;
    }
    return 0;
}

/* y[n,o,oy,ox] = sum_{c,i,j} w'[o,c,i,j] * xpad[n,c,oy*stride + i, ox*stride + j];  w' = w (flip_weight != 0, cross-correlation) or w mirrored. */
int orc_conv2d(const float* x, const float* w, float* y, int N, int Cin, int H, int W, int Cout, int kh, int kw,
               int stride, int pad_y, int pad_x, int flip_weight) {
    const int outH = (H + 2 * pad_y - kh) / stride + 1, outW = (W + 2 * pad_x - kw) / stride + 1;
    if (outH < 1 || outW < 1) return 1;
#pragma omp parallel for collapse(2)
    for (int n = 0; n < N; n++)
        for (int o = 0; o < Cout; o++)
            for (int oy = 0; oy < outH; oy++)
                for (int ox = 0; ox < outW; ox++) {
                    double acc = 0.0;
                    for (int c = 0; c < Cin; c++)
                        for (int i = 0; i < kh; i++) {
                            const int iy = oy * stride + i - pad_y;
                            if (iy < 0 || iy >= H) continue;
                            for (int j = 0; j < kw; j++) {
                                const int ix = ox * stride + j - pad_x;
                                if (ix < 0 || ix >= W) continue;
                                const int wi = flip_weight ? i : kh - 1 - i, wj = flip_weight ? j : kw - 1 - j;
                                acc += (double)w[(((size_t)o * Cin + c) * kh + wi) * kw + wj] * (double)x[(((size_t)n * Cin + c) * H + iy) * W + ix];
                            }
                        }
                    y[(((size_t)n * Cout + o) * outH + oy) * outW + ox] = (float)acc;
                }
    return 0;
}
