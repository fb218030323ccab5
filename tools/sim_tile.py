"""Host-side model of tile_kernel.cuh: checks the stage algebra against numpy and counts
shared-memory bank conflicts per access for a given (elt_bytes, L, R, W, variant)."""
import numpy as np, sys, itertools

def ilog2(v): return v.bit_length() - 1

def fold(idx, bits, total):
    f = 0; s = bits
    while s < total:
        f ^= idx >> s; s += bits
    return f & ((1 << bits) - 1)

def simulate(L, R, W, elt=16, load_row=False, store_row=False, check=True):
    LOG_L, LOG_R, LOG_W = ilog2(L), ilog2(R), ilog2(W)
    S = (LOG_L + LOG_R - 1) // LOG_R
    R_LAST = L >> (LOG_R * (S - 1))
    T_LINE = L // R; LOG_TL = ilog2(T_LINE)
    THREADS = T_LINE * W
    SB = 3 if elt == 16 else 4
    IDX_BITS = LOG_L + LOG_W
    lanes_per_phase = 128 // elt
    rng = np.random.default_rng(0)
    x = rng.standard_normal((L, W)) + 1j * rng.standard_normal((L, W))
    tw = np.exp(-2j * np.pi * np.arange(L) / L)
    sm = np.zeros(L * W, complex)
    conflicts = {}
    def swz(idx): return idx ^ fold(idx, SB, IDX_BITS)
    def record(name, addrs_by_thread):
        # addrs_by_thread: array [THREADS] of element addresses for ONE instruction
        worst = 1
        for w0 in range(0, THREADS, 32):
            warp = addrs_by_thread[w0:w0 + 32]
            for ph in range(0, len(warp), lanes_per_phase):
                grp = warp[ph:ph + lanes_per_phase]
                banks = {}
                for a in grp:
                    banks.setdefault(a % lanes_per_phase, set()).add(a)
                worst = max(worst, max(len(s) for s in banks.values()))
        conflicts[name] = max(conflicts.get(name, 1), worst)
    ts = np.arange(THREADS)
    w_col, u_col = ts & (W - 1), ts >> LOG_W
    u_row, w_row = ts & (T_LINE - 1), ts >> LOG_TL
    w1, u1 = (w_row, u_row) if load_row else (w_col, u_col)
    v = np.zeros((THREADS, R), complex)
    for d in range(R):
        v[:, d] = x[u1 + d * T_LINE, w1]
    out = np.zeros((L, W), complex)
    if S > 1:
        v = np.fft.fft(v, axis=1)
        for d in range(1, R): v[:, d] *= tw[d * u1]
        base = (u1 << LOG_W) | w1
        for d in range(R):
            a = swz(base | ((d * T_LINE) << LOG_W)); sm[a] = v[:, d]; record("st1", a)
        for s in range(2, S):
            log_ms = LOG_L - LOG_R * s; ms = 1 << log_ms
            lo = u_col & (ms - 1); hi = u_col >> log_ms
            pos = (hi << (log_ms + LOG_R)) | lo
            base = (pos << LOG_W) | w_col
            for d in range(R):
                a = swz(base | ((d << log_ms) << LOG_W)); v[:, d] = sm[a]; record(f"ld{s}", a)
            v = np.fft.fft(v, axis=1)
            tstep = lo << (LOG_R * (s - 1))
            for d in range(1, R): v[:, d] *= tw[d * tstep]
            for d in range(R):
                a = swz(base | ((d << log_ms) << LOG_W)); sm[a] = v[:, d]
        B = R // R_LAST
        wl, ul = (w_row, u_row) if store_row else (w_col, u_col)
        for b in range(B):
            jp = ul + b * T_LINE
            pos = np.zeros_like(jp)
            for i in range(1, S):
                ki = (jp >> (LOG_R * (i - 1))) & (R - 1)
                pos |= ki << (LOG_L - LOG_R * i)
            base = (pos << LOG_W) | wl
            for n in range(R_LAST):
                a = swz(base | (n << LOG_W)); v[:, b * R_LAST + n] = sm[a]; record("ldL", a)
            v[:, b * R_LAST:(b + 1) * R_LAST] = np.fft.fft(v[:, b * R_LAST:(b + 1) * R_LAST], axis=1)
            for q in range(R_LAST):
                out[jp + q * (L // R_LAST), wl] = v[:, b * R_LAST + q]
    else:
        v = np.fft.fft(v, axis=1)
        wl = w_row if store_row else w_col
        for q in range(R): out[q, wl] = v[:, q]
    err = None
    if check:
        ref = np.fft.fft(x, axis=0)
        err = np.abs(out - ref).max() / np.abs(ref).max()
    return dict(S=S, R_LAST=R_LAST, threads=THREADS, smem=L * W * elt, err=err, conflicts=conflicts)

if __name__ == "__main__":
    for elt, cfgs in [(16, [(2,2,8),(4,4,8),(8,8,8),(16,8,8),(32,8,8),(64,8,8),(128,8,8),(256,8,8),(512,8,8),(512,8,4),(1024,16,8),(1024,16,4),(2048,16,4),(2048,16,2),(4096,16,2),(4096,16,1),(8192,16,1),(1024,8,4)]),
                      (8, [(16,16,16),(32,8,16),(64,8,16),(128,16,16),(256,16,16),(512,16,16),(1024,16,16),(2048,16,8),(4096,16,4),(8192,16,2),(16384,16,1),(512,8,16)])]:
        for (L, R, W) in cfgs:
            for lr, sr in [(False, False), (True, True), (True, False)]:
                r = simulate(L, R, W, elt, lr, sr)
                bad = {k: v for k, v in r["conflicts"].items() if v > 1}
                print(f"elt={elt} L={L} R={R} W={W} {'R' if lr else 'C'}{'R' if sr else 'C'} S={r['S']} rl={r['R_LAST']} thr={r['threads']} smem={r['smem']//1024}K err={r['err']:.1e} conflicts={bad}")
