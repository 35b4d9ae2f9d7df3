"""Development aid: summarise the per-phase timeline the tiled kernel dumps with PBD_B200_TRACE=<file> (ns)."""
import sys, numpy as np
d = np.fromfile(sys.argv[1], dtype=np.uint64)
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 148
d = d.reshape(-1, nt, 4).astype(np.int64)
ok = (d[:, :, 3] > 0).all(axis=1)
d = d[ok]
print("phases traced", len(d))
rel = d[1:, :, 3].max(axis=1)  # release time of phase p (last CTA)
prev = d[:-1]
cur = d[1:]
print("phase length (release to release), us: median %.2f mean %.2f" % (np.median(np.diff(d[:, :, 3].max(axis=1))) / 1e3, np.mean(np.diff(d[:, :, 3].max(axis=1))) / 1e3))
A = cur[:, :, 0] - prev[:, :, 3]      # end-of-colour sync -> all spanning items of the CTA collected
at = cur[:, :, 1] - cur[:, :, 0]      # fence + atomic
W = cur[:, :, 2] - cur[:, :, 1]       # spin until every CTA arrived
B = cur[:, :, 3] - cur[:, :, 2]       # release -> workers done with the private items too
for name, x in (("A collected", A), ("fence+atomic", at), ("spin", W), ("B tail", B)):
    print("%-16s per CTA: median %.2f us, mean %.2f, p95 %.2f, max-over-CTAs median %.2f" % (name, np.median(x) / 1e3, x.mean() / 1e3, np.percentile(x, 95) / 1e3, np.median(x.max(axis=1)) / 1e3))
for p in range(1, min(len(d), 8)):
    print("phase %3d: len %.2f  A(max) %.2f  spin(min) %.2f  Btail(max) %.2f" % (p, (d[p, :, 3].max() - d[p - 1, :, 3].max()) / 1e3, A[p - 1].max() / 1e3, W[p - 1].min() / 1e3, B[p - 1].max() / 1e3))

if len(sys.argv) > 3:  # worker trace: [0] phase start, [1] operands there, [2] spanning done, [3] private done
    w0 = d[:, :, 1] - d[:, :, 0]; w1 = d[:, :, 2] - d[:, :, 1]; w2 = d[:, :, 3] - d[:, :, 2]
    gap = d[1:, :, 0] - d[:-1, :, 3]
    for name, x in (("wait operands", w0), ("spanning part", w1), ("private part", w2), ("colour end -> next start", gap)):
        print("worker %-24s median %.2f us, mean %.2f, p95 %.2f" % (name, np.median(x) / 1e3, x.mean() / 1e3, np.percentile(x, 95) / 1e3))
if len(sys.argv) > 3:
    start = d[:, :, 0]
    for p in (5, 12, 25, 26):
        if p + 1 < len(d):
            print("phase %d: start skew %.2f us, phase len %.2f us" % (p, (start[p].max() - start[p].min()) / 1e3, (np.median(start[p + 1]) - np.median(start[p])) / 1e3))
