"""numpy restatement of GGML_OP_ROPE (test infrastructure): pinned to the reference CPU implementation by tests/test_rope_pin.py, used by
tests/test_gpu_ops.py as the checker of b200_op_rope"""
import numpy as np


def rope_numpy(x, pos, n_dims, mode, n_orig_ctx, freq_base, freq_scale, ext_factor, attn_factor, beta_fast, beta_slow, xpos_base=0.0, xpos_down=False):
    """ggml_compute_forward_rope_f32 / _f16 (src/ggml.c:13775, :13953), forward, fp32 arithmetic kept fp32 step by step.  x: [b][tokens][heads][ne0]"""
    f = np.float32
    B, T, H, ne0 = x.shape
    out = x.astype(np.float32).copy()
    theta_scale = f(np.float64(f(freq_base)) ** np.float64(f(-2.0) / f(n_dims)))     # glibc powf / cosf / sinf / logf are correctly rounded: double then round
    corr = [f(0), f(0)]
    if ext_factor != 0.0:
        def corr_dim(n_rot):
            return f(n_dims) * f(np.log(np.float64(f(n_orig_ctx) / (f(n_rot) * f(2) * f(np.pi))))) / (f(2) * f(np.log(np.float64(f(freq_base)))))
        corr = [max(f(0), np.floor(corr_dim(beta_fast))), min(f(n_dims - 1), np.ceil(corr_dim(beta_slow)))]

    def yarn(theta_extrap, i0):
        theta_interp = f(freq_scale) * theta_extrap
        theta, mscale = theta_interp, f(attn_factor)
        if ext_factor != 0.0:
            y = (f(i0 // 2) - corr[0]) / max(f(0.001), corr[1] - corr[0])
            ramp_mix = (f(1) - min(f(1), max(f(0), y))) * f(ext_factor)
            theta = theta_interp * (f(1) - ramp_mix) + theta_extrap * ramp_mix
            mscale = mscale * (f(1) + f(0.1) * f(np.log(np.float64(f(1) / f(freq_scale)))))
        return f(f(np.cos(np.float64(theta))) * mscale), f(f(np.sin(np.float64(theta))) * mscale)

    xf = x.astype(np.float32)
    for t in range(T):
        p = int(pos[t])
        if not (mode & 2):
            theta = f(p)
            for i0 in range(0, ne0, 2):
                c, s = yarn(theta, i0)
                zeta = f(1)
                if xpos_base != 0.0 and x.dtype == np.float32:
                    zeta = f(np.float64((f(i0) + f(0.4) * f(ne0)) / (f(1.4) * f(ne0))) ** np.float64(f(p) / f(xpos_base)))
                    if xpos_down:
                        zeta = f(1) / zeta
                x0, x1 = xf[:, t, :, i0], xf[:, t, :, i0 + 1]
                out[:, t, :, i0] = x0 * c * zeta - x1 * s * zeta
                out[:, t, :, i0 + 1] = x0 * s * zeta + x1 * c * zeta
                theta = f(theta * theta_scale)
        else:
            theta = f(f(p) * f(freq_scale))
            for ic in range(0, n_dims, 2):
                c, s = yarn(theta, 0)
                theta = f(theta * theta_scale)
                i0, i1 = ic // 2, ic // 2 + n_dims // 2
                x0, x1 = xf[:, t, :, i0], xf[:, t, :, i1]
                out[:, t, :, i0] = x0 * c - x1 * s
                out[:, t, :, i1] = x0 * s + x1 * c
    return out.astype(x.dtype)
