"""Synthetic spatial networks for configs C2-C5 in the engine's exchange format:
post-sorted CSR (rows = target neuron, in-row ascending presynaptic ID), the iteration order of the
reference's Neuron::inSynapses map (NeuCor.h:212, NeuCor.cpp:690).

Recipe (the reference has no large-network builder; NeuCor(int) is O(N^2) and fixed-degree):
  * positions uniform in a cube of side (N/8)^(1/3) — the reference's spawn density of 8 neurons
    per unit volume (NeuCor.cpp:155);
  * every neuron draws K distinct presynaptic partners uniformly from the neurons within radius R
    of it (R chosen so the ball holds ~2K neurons), rejecting lengths < MIN_LENGTH so that every
    synaptic delay 2*length exceeds the step dt = 0.0625 ms; neurons close to the cube's faces
    that see fewer than K candidates take all of them;
  * weights U(0.2, 1), 20 % negated, inhibitory flag = sign (Synapse ctor, NeuCor.cpp:471-475);
  * G = N/250 input firers on a jittered grid, radius 0.8 (as the STANDARD preset, main.cpp:87),
    `near` lists in ascending neuron ID (NeuCor.cpp:319-323).
"""
import numpy as np

MIN_LENGTH = 0.04
DENSITY = 8.0


def _dist32(a, b):
    """coord3::getDist (NeuCor.h:16-18) in float32: sqrtf(dx*dx + dy*dy + dz*dz), left to right."""
    d = a[:, None, :] - b[None, :, :]
    d2 = d[..., 0] * d[..., 0]
    d2 = d2 + d[..., 1] * d[..., 1]
    d2 = d2 + d[..., 2] * d[..., 2]
    return np.sqrt(d2)


def radius_for(K):
    """Ball radius that holds ~2K neurons at the reference density."""
    return float((2.0 * K / (DENSITY * 4.0 / 3.0 * np.pi)) ** (1.0 / 3.0))


def synthetic_network(N, K, seed=1, R=None, inputs_per_neuron=1.0 / 250.0, input_radius=0.8):
    rng = np.random.default_rng(seed)
    N, K = int(N), int(K)
    L = (N / DENSITY) ** (1.0 / 3.0)
    R = radius_for(K) if R is None else float(R)
    pos = (rng.random((N, 3)) * L).astype(np.float32)
    nc = max(1, int(np.floor(L / R)))
    cs = L / nc
    cell = np.minimum((pos / cs).astype(np.int64), nc - 1)
    cid = (cell[:, 0] * nc + cell[:, 1]) * nc + cell[:, 2]
    order = np.argsort(cid, kind="stable")
    starts = np.searchsorted(cid[order], np.arange(nc ** 3 + 1))

    def members(cx, cy, cz):
        out = []
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dz in (-1, 0, 1):
                    x, y, z = cx + dx, cy + dy, cz + dz
                    if 0 <= x < nc and 0 <= y < nc and 0 <= z < nc:
                        c = (x * nc + y) * nc + z
                        out.append(order[starts[c]:starts[c + 1]])
        return np.concatenate(out) if out else np.zeros(0, np.int64)

    rowlen = np.zeros(N, np.int64)
    rows_pre = [None] * N
    rows_len = [None] * N
    for cx in range(nc):
        for cy in range(nc):
            for cz in range(nc):
                c = (cx * nc + cy) * nc + cz
                mine = order[starts[c]:starts[c + 1]]
                if len(mine) == 0:
                    continue
                cand = members(cx, cy, cz)
                d = _dist32(pos[mine], pos[cand])
                ok = (d < R) & (d >= MIN_LENGTH) & (mine[:, None] != cand[None, :])
                keys = rng.random(d.shape)
                keys[~ok] = 2.0
                take = min(K, d.shape[1])
                sel = np.argpartition(keys, take - 1, axis=1)[:, :take] if take < d.shape[1] else np.tile(np.arange(d.shape[1]), (len(mine), 1))
                for i, q in enumerate(mine):
                    s = sel[i]
                    s = s[keys[i, s] < 1.5]
                    p = cand[s]
                    o = np.argsort(p)
                    rows_pre[q] = p[o].astype(np.uint32)
                    rows_len[q] = d[i, s][o].astype(np.float32)
                    rowlen[q] = len(p)
    rowptr = np.zeros(N + 1, np.uint64)
    rowptr[1:] = np.cumsum(rowlen)
    S = int(rowptr[N])
    pre = np.concatenate([r for r in rows_pre if r is not None and len(r)]) if S else np.zeros(0, np.uint32)
    length = np.concatenate([r for r in rows_len if r is not None and len(r)]) if S else np.zeros(0, np.float32)
    w = (rng.random(S).astype(np.float32) * np.float32(0.8) + np.float32(0.2)).astype(np.float32)
    neg = rng.random(S) < 0.2
    w[neg] = -w[neg]
    net = dict(N=N, S=S, rowptr=rowptr, pre=pre.astype(np.uint32), weight=w, length=length,
               flag=(w < 0).astype(np.uint8), positions=pos)
    # input firers on a jittered grid
    G = max(1, int(round(N * inputs_per_neuron)))
    gpos = (rng.random((G, 3)) * L).astype(np.float32)
    near = []
    for g in range(G):
        d = _dist32(gpos[g:g + 1], pos)[0]
        near.append(np.nonzero(d < np.float32(input_radius))[0].astype(np.uint32))
    net["inputs"] = dict(G=G, positions=gpos, radius=np.full(G, input_radius, np.float32), near=near)
    return net


def uniform_random_network(N, K, seed=1, max_length=None):
    """Cheap stand-in used for throughput runs at sizes the spatial builder cannot reach from numpy:
    K distinct uniformly random presynaptic partners per neuron, lengths drawn from the distance
    distribution of uniform points in a ball of radius R (pdf ~ r^2) truncated to >= MIN_LENGTH."""
    rng = np.random.default_rng(seed)
    N, K = int(N), int(K)
    R = radius_for(K) if max_length is None else float(max_length)
    pre = np.empty((N, K), np.uint32)
    for q0 in range(0, N, 4096):
        q1 = min(N, q0 + 4096)
        # K distinct draws per row: sample with replacement, redraw duplicates/self until clean
        p = rng.integers(0, N, size=(q1 - q0, K), dtype=np.int64)
        for _ in range(64):
            p.sort(axis=1)
            bad = np.zeros_like(p, bool)
            bad[:, 1:] = p[:, 1:] == p[:, :-1]
            bad |= p == np.arange(q0, q1)[:, None]
            nb = int(bad.sum())
            if nb == 0:
                break
            p[bad] = rng.integers(0, N, size=nb, dtype=np.int64)
        p.sort(axis=1)
        pre[q0:q1] = p
    S = N * K
    length = (R * rng.random(S) ** (1.0 / 3.0)).astype(np.float32)
    length = np.maximum(length, np.float32(MIN_LENGTH))
    w = (rng.random(S).astype(np.float32) * np.float32(0.8) + np.float32(0.2)).astype(np.float32)
    neg = rng.random(S) < 0.2
    w[neg] = -w[neg]
    rowptr = (np.arange(N + 1, dtype=np.uint64) * np.uint64(K))
    G = max(1, N // 250)
    near = [np.sort(rng.choice(N, size=min(N, 17), replace=False)).astype(np.uint32) for _ in range(G)]
    return dict(N=N, S=S, rowptr=rowptr, pre=pre.reshape(-1), weight=w, length=length,
                flag=(w < 0).astype(np.uint8), positions=None,
                inputs=dict(G=G, positions=None, radius=np.full(G, 0.8, np.float32), near=near))


def stratified_network_torch(N, K, device, seed=1, chunk_rows=32768, weight_scale=1.0):
    """Whole network of stratified_shard_torch (one shard holding every row)."""
    return stratified_shard_torch(N, K, 0, N, device, seed=seed, chunk_rows=chunk_rows, weight_scale=weight_scale)


def stratified_shard_torch(N, K, row0, n_rows, device, seed=1, chunk_rows=32768, weight_scale=1.0, near_size=17):
    """Rows [row0, row0+n_rows) of the C3-scale stand-in, built directly in device memory with torch (plumbing only):
    every neuron gets K distinct presynaptic partners, one drawn uniformly from each of K equal strata of the GLOBAL ID
    range (so rows are sorted by construction and in-degree is exactly K), lengths from the distance distribution of
    uniform points in a ball of radius radius_for(K) (pdf ~ r^2) truncated to >= MIN_LENGTH, weights
    U(0.2, 1) * weight_scale with 20 % negated.  Returns device tensors (local rowptr int64 from 0, pre int32, weight
    f32, length f32, flag uint8), the host `near` lists of the WHOLE network's input firers (identical on every shard)
    and `min_delay` of this shard."""
    import torch
    gen = torch.Generator(device=device)
    gen.manual_seed(seed * 1000003 + row0)
    N, K, row0, n_rows = int(N), int(K), int(row0), int(n_rows)
    S = n_rows * K
    R = radius_for(K)
    pre = torch.empty(S, dtype=torch.int32, device=device)
    weight = torch.empty(S, dtype=torch.float32, device=device)
    length = torch.empty(S, dtype=torch.float32, device=device)
    bounds = (torch.arange(K + 1, device=device, dtype=torch.int64) * N) // K
    lo, width = bounds[:-1], (bounds[1:] - bounds[:-1])
    for q0 in range(0, n_rows, chunk_rows):
        q1 = min(n_rows, q0 + chunk_rows)
        n = q1 - q0
        u = torch.rand((n, K), device=device, generator=gen, dtype=torch.float64)
        off = torch.minimum((u * width).to(torch.int64), width - 1)
        p = lo + off
        rows = torch.arange(row0 + q0, row0 + q1, device=device, dtype=torch.int64)[:, None]
        clash = p == rows
        p = torch.where(clash, lo + (off + 1) % width, p)
        pre[q0 * K:q1 * K] = p.reshape(-1).to(torch.int32)
        ul = torch.rand(n * K, device=device, generator=gen)
        length[q0 * K:q1 * K] = torch.clamp(R * ul.pow(1.0 / 3.0), min=MIN_LENGTH)
        w = (torch.rand(n * K, device=device, generator=gen) * 0.8 + 0.2) * float(weight_scale)
        neg = torch.rand(n * K, device=device, generator=gen) < 0.2
        weight[q0 * K:q1 * K] = torch.where(neg, -w, w)
        del u, off, p, clash, ul, w, neg
    flag = (weight < 0).to(torch.uint8)
    rowptr = torch.arange(n_rows + 1, device=device, dtype=torch.int64) * K
    rng = np.random.default_rng(seed)
    G = max(1, N // 250)
    # `near_size` distinct random neurons per firer, ascending (one vectorised draw, de-duplicated per firer: O(G * near_size))
    draws = np.sort(rng.integers(0, N, size=(G, min(N, near_size))), axis=1)
    keep = np.ones(draws.shape, bool)
    keep[:, 1:] = draws[:, 1:] != draws[:, :-1]
    near = [draws[g][keep[g]].astype(np.uint32) for g in range(G)]
    return dict(N=N, S=S, row0=row0, n_rows=n_rows, rowptr=rowptr, pre=pre, weight=weight, length=length, flag=flag,
                min_delay=float(length.min().item()) * 2.0 if S else float("inf"),
                inputs=dict(G=G, positions=None, radius=np.full(G, 0.8, np.float32), near=near))


def spatial_shard_torch(N, K, row0, n_rows, device, seed=1, weight_scale=1.0, time_budget_s=None, pad_rows=65536):
    """Rows [row0, row0+n_rows) of the SURVEY.md section 8(d) recipe at C2-C4 scale, built in device memory with torch
    (plumbing only; runs unchanged on CPU tensors, which is how tests/test_networks.py checks it):
      * positions of ALL N neurons uniform in a cube of side (N/8)^(1/3) — identical on every rank (same seed);
      * every neuron of this shard draws min(K, candidates) distinct presynaptic partners uniformly from the neurons within
        radius_for(K) of it, lengths < MIN_LENGTH rejected, no self-synapse; lengths are coord3::getDist in float32
        (NeuCor.h:16-18), rows ascend in presynaptic ID; neurons near the cube's faces that see fewer than K candidates take
        all of them (ragged rows);
      * weights U(0.2, 1) * weight_scale, 20 % negated, flag = sign.
    A unit-cell grid of side >= R bounds the candidates to the 27 surrounding cells; within a cell block the K partners are
    the K smallest of one uniform key per candidate pair (topk).  Returns the same dict as stratified_shard_torch plus
    `positions` (host float32 [N, 3]) — input firers get their `near` lists from the host class's grid (radius 0.8).
    Raises TimeoutError when time_budget_s is exceeded (callers fall back to the stratified stand-in)."""
    import time
    import torch
    t_start = time.perf_counter()
    N, K, row0, n_rows = int(N), int(K), int(row0), int(n_rows)
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed * 1000003)
    L = (N / DENSITY) ** (1.0 / 3.0)
    R = radius_for(K)
    pos = torch.rand((N, 3), generator=gen, device=dev, dtype=torch.float32) * L
    gen.manual_seed(seed * 1000003 + 17 + row0)  # what follows is per shard
    nc = max(1, int(np.floor(L / R)))
    cs = L / nc
    cell3 = torch.clamp((pos / cs).to(torch.int64), max=nc - 1)
    cid = (cell3[:, 0] * nc + cell3[:, 1]) * nc + cell3[:, 2]
    order = torch.argsort(cid, stable=True)
    counts = torch.bincount(cid, minlength=nc ** 3)
    starts = torch.zeros(nc ** 3 + 1, dtype=torch.int64, device=dev)
    starts[1:] = torch.cumsum(counts, 0)
    starts_h = starts.cpu().numpy()
    mine_mask = torch.zeros(N, dtype=torch.bool, device=dev)
    mine_mask[row0:row0 + n_rows] = True
    # padded per-row results, compacted at the end (rows are ragged near the faces)
    pre_pad = torch.empty((n_rows, K), dtype=torch.int32, device=dev)
    len_pad = torch.empty((n_rows, K), dtype=torch.float32, device=dev)
    cnt = torch.zeros(n_rows, dtype=torch.int64, device=dev)
    Rf, minlen = np.float32(R), np.float32(MIN_LENGTH)
    for cx in range(nc):
        for cy in range(nc):
            if time_budget_s is not None and time.perf_counter() - t_start > time_budget_s:
                raise TimeoutError("spatial_shard_torch: %.0f s budget exceeded" % time_budget_s)
            for cz in range(nc):
                c = (cx * nc + cy) * nc + cz
                members = order[starts_h[c]:starts_h[c + 1]]
                mine = members[mine_mask[members]]
                if mine.numel() == 0:
                    continue
                parts = []
                for dx in (-1, 0, 1):
                    for dy in (-1, 0, 1):
                        for dz in (-1, 0, 1):
                            x, y, z = cx + dx, cy + dy, cz + dz
                            if 0 <= x < nc and 0 <= y < nc and 0 <= z < nc:
                                q = (x * nc + y) * nc + z
                                parts.append(order[starts_h[q]:starts_h[q + 1]])
                cand = torch.sort(torch.cat(parts)).values
                pm, pc = pos[mine], pos[cand]
                d0 = pm[:, None, 0] - pc[None, :, 0]
                d2 = d0 * d0
                d1 = pm[:, None, 1] - pc[None, :, 1]
                d2 = d2 + d1 * d1
                d1 = pm[:, None, 2] - pc[None, :, 2]
                d2 = d2 + d1 * d1
                d = torch.sqrt(d2.double()).float()  # = correctly rounded sqrtf (torch's own float32 sqrt is 1 ulp off on some CPUs)
                ok = (d < Rf) & (d >= minlen) & (mine[:, None] != cand[None, :])
                keys = torch.rand(d.shape, generator=gen, device=dev, dtype=torch.float32)
                keys = torch.where(ok, keys, torch.full_like(keys, 2.0))
                kk = min(K, cand.numel())
                vals, idx = torch.topk(keys, kk, dim=1, largest=False)
                valid = vals < 1.5
                idx = torch.where(valid, idx, torch.full_like(idx, cand.numel()))
                idx = torch.sort(idx, dim=1).values  # ascending candidate position = ascending presynaptic ID; rejected ones last
                n_ok = valid.sum(1)
                safe = torch.clamp(idx, max=cand.numel() - 1)
                rows = mine - row0
                pre_pad[rows, :kk] = cand[safe].to(torch.int32)
                len_pad[rows, :kk] = torch.gather(d, 1, safe)
                cnt[rows] = n_ok
                del d, d0, d1, d2, ok, keys, vals, idx, safe, valid
    rowptr = torch.zeros(n_rows + 1, dtype=torch.int64, device=dev)
    rowptr[1:] = torch.cumsum(cnt, 0)
    S = int(rowptr[-1].item())
    pre = torch.empty(S, dtype=torch.int32, device=dev)
    length = torch.empty(S, dtype=torch.float32, device=dev)
    cols = torch.arange(K, device=dev)
    for a in range(0, n_rows, pad_rows):  # compact the padded rows chunk by chunk
        b = min(n_rows, a + pad_rows)
        m = cols[None, :] < cnt[a:b, None]
        lo, hi = int(rowptr[a].item()), int(rowptr[b].item())
        pre[lo:hi] = pre_pad[a:b][m]
        length[lo:hi] = len_pad[a:b][m]
    del pre_pad, len_pad
    weight = torch.empty(S, dtype=torch.float32, device=dev)
    for a in range(0, S, 1 << 26):
        b = min(S, a + (1 << 26))
        w = (torch.rand(b - a, device=dev, generator=gen) * 0.8 + 0.2) * float(weight_scale)
        neg = torch.rand(b - a, device=dev, generator=gen) < 0.2
        weight[a:b] = torch.where(neg, -w, w)
    flag = (weight < 0).to(torch.uint8)
    rng = np.random.default_rng(seed)
    G = max(1, N // 250)
    gpos = (rng.random((G, 3)) * L).astype(np.float32)
    return dict(N=N, S=S, row0=row0, n_rows=n_rows, rowptr=rowptr, pre=pre, weight=weight, length=length, flag=flag,
                min_delay=float(length.min().item()) * 2.0 if S else float("inf"), positions=pos.cpu().numpy(),
                inputs=dict(G=G, positions=gpos, radius=np.full(G, 0.8, np.float32), near=None))
