"""Enumerates every shared-memory index fk_stft (quadrs_b200/csrc/qd_stft.cu) forms and checks that the padded
index is the padded base plus a compile-time offset: pad_idx(base + q*i) == pad_idx(base) + pad_off(q, i)."""


def pad_idx(i):
    return i + (i >> 4) + (i >> 8)


def pad_off(q, i):
    return q * i + ((q * i) >> 4) + ((q * i) >> 8)


def check():
    bad = 0
    for logw in range(5, 13):
        W = 1 << logw
        odd = logw & 1
        first = 3 if odd else 4
        rest = logw - first
        n16, n4 = rest // 4, (rest % 4) // 2
        tw = W // 16
        q = 1 << first
        n = 8 if odd else 16
        for g in range(W // n):  # pass-1 stores of one group
            bad += sum(pad_idx(n * g + i) != pad_idx(n * g) + i for i in range(n))
        for _ in range(n16):
            for lt in range(tw):
                base = (lt // q) * 16 * q + (lt & (q - 1))
                bad += sum(pad_idx(base + q * i) != pad_idx(base) + pad_off(q, i) for i in range(16))
            q *= 16
        if n4:
            for gid in range(4 * tw):
                base = (gid // q) * 4 * q + (gid & (q - 1))
                bad += sum(pad_idx(base + q * i) != pad_idx(base) + pad_off(q, i) for i in range(4))
    return bad


if __name__ == "__main__":
    b = check()
    print("mismatches:", b)
    raise SystemExit(1 if b else 0)
