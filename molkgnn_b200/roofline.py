"""Algorithmic (compulsory) HBM bytes and FLOPs of the conv stack -- SURVEY.md 8(d), BASELINE.md 3.

Per layer, each tensor touched once, int32 topology, uint8 arg-max, kernel parameters excluded (batch independent):
  fwd(F,K) = 4NF [x] + 4E*Fe [bond attrs] + 4E + 4N [CSR] + 4NK [h_out] + sum_d n_d L_d [argmax] (+12N [p] last layer)
  bwd(F,K) = 4NK [grad_h] + 4NF [x] + sum_d n_d L_d + 4E*Fe + 4E + 4N + 4NF [grad_x]
"""


def layer_bytes(N, E, n, F, L, Fe=7, last=False, training=True):
    K = sum(L)
    am = sum(nd * ld for nd, ld in zip(n, L))
    fwd = 4 * N * F + 4 * E * Fe + 4 * E + 4 * N + 4 * N * K + (am if training else 0) + (12 * N if last else 0)
    bwd = 4 * N * K + 4 * N * F + am + 4 * E * Fe + 4 * E + 4 * N + 4 * N * F
    return fwd, bwd


def stack_bytes(N, E, n, x_dim, L1, LN, num_layers, Fe=7, training=True):
    """training=False: forward-only sweep (BASELINE configs[4]) -- no arg-max store"""
    fwd = bwd = 0
    F = x_dim
    for i in range(num_layers):
        L = L1 if i == 0 else LN
        f, b = layer_bytes(N, E, n, F, L, Fe, last=(i == num_layers - 1), training=training)
        fwd += f
        bwd += b
        F = sum(L)
    return fwd, bwd


def layer_flops(n, F, L, Fe=7, E=0):
    """fwd FLOPs: 2F sum n_d L_d (d^2+1) + 2Fe sum n_d L_d d^2 + E K;  bwd MACs ~ F sum n_d L_d (3d+3)."""
    K = sum(L)
    fwd = sum(2 * F * nd * ld * (d * d + 1) + 2 * Fe * nd * ld * d * d for d, (nd, ld) in enumerate(zip(n, L), 1))
    fwd += E * K
    bwd = sum(2 * F * nd * ld * (3 * d + 3) for d, (nd, ld) in enumerate(zip(n, L), 1))
    return fwd, bwd


def stack_flops(E, n, x_dim, L1, LN, num_layers, Fe=7):
    fwd = bwd = 0
    F = x_dim
    for i in range(num_layers):
        L = L1 if i == 0 else LN
        f, b = layer_flops(n, F, L, Fe, E)
        fwd += f
        bwd += b
        F = sum(L)
    return fwd, bwd


def kernel_bytes_per_step(kernel, N, E, n, x_dim, L1, LN, num_layers, Fe=7):
    """Share of the per-layer algorithmic bytes (formulas above) that one kernel class is responsible for, summed
    over the layers of one step.  Intermediates of the unfused design (compact scores, coefficients, per-CTA partial
    sums, the second read of x in the input-gradient kernel) are NOT counted: they lower the achieved fraction."""
    tot = 0
    F = x_dim
    for i in range(num_layers):
        L = L1 if i == 0 else LN
        K = sum(L)
        am = sum(nd * ld for nd, ld in zip(n, L))
        topo = 4 * E + 4 * N
        if kernel == "stack_fwd_fused":      # layer-fused forward: the whole forward formula (per-layer contract)
            tot += layer_bytes(N, E, n, F, L, Fe, last=(i == num_layers - 1))[0]
        elif kernel == "conv_fwd":
            tot += 4 * N * F + 4 * E * Fe + topo + am + (12 * N if i == num_layers - 1 else 0)
        elif kernel == "propagate_fwd":
            tot += 4 * N * K
        elif kernel == "bwd_w":
            tot += 4 * N * K + 4 * N * F + am + 4 * E * Fe + topo
        elif kernel == "bwd_x":
            tot += 4 * N * F
        elif kernel == "coef_bond":          # tile backward, pre-pass: grad_h gather, arg-max, bond rows
            tot += 4 * N * K + am + 4 * E * Fe + topo
        elif kernel == "conv_bwd_tile":      # tile backward, main kernel: x (as fp16 images) in, grad_x out
            tot += 8 * N * F + topo
        elif kernel in ("bucket_build", "bucket_assign"):
            tot += (16 * E + 12 * N + 4 * E * Fe) if i == 0 else 0
        else:
            tot += 0
        F = K
    return tot
