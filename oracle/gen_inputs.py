"""TEST/BENCH INFRASTRUCTURE -- seeded procedural inputs of SURVEY.md section 8(d) (no KITTI / PASCAL / SUN data
exists here).  numpy only, so bench.py can use it for both arms without touching the oracle's arithmetic."""
import numpy as np


def kitti_k(h, w):
    return np.array([[721.54 * w / 1242.0, 0, 609.56 * w / 1242.0],
                     [0, 721.54 * h / 375.0, 172.85 * h / 375.0], [0, 0, 1.0]], dtype=np.float32)


def synth_k(h, w):
    return np.array([[w, 0, w / 2.0], [0, h, h / 2.0], [0, 0, 1.0]], dtype=np.float32)   # syntheticPlanes/data.py:548-557


def _rot(ax, ay, az):
    cx, sx, cy, sy, cz, sz = np.cos(ax), np.sin(ax), np.cos(ay), np.sin(ay), np.cos(az), np.sin(az)
    rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return (rz @ ry @ rx).astype(np.float32)


def band_limited(rs, shape_hw, channels, n_waves=8):
    """Sum of random sinusoids per channel, scaled to [0,1]."""
    h, w = shape_hw
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    out = np.zeros((h, w, channels), np.float32)
    for c in range(channels):
        for _ in range(n_waves):
            fx, fy = rs.uniform(-0.15, 0.15, 2)
            out[..., c] += rs.uniform(0.3, 1.0) * np.sin(fx * xx + fy * yy + rs.uniform(0, 6.28))
    out -= out.min()
    return out / max(out.max(), 1e-6)


def scene(L, B, H, W, cam, seed, max_disp):
    """Returns dict of float32 arrays: tex [L,B,H,W,3], mask [L,B,H,W,1], disp [L,B,H,W,1], k_s,k_t,rot [B,3,3],
    t [B,3,1].  cam: 'identity' (config 1), 'kitti' (configs 2,4,5), 'synth' (config 3)."""
    rs = np.random.RandomState(seed)
    tex = np.stack([np.stack([band_limited(rs, (H, W), 3) for _ in range(B)]) for _ in range(L)]).astype(np.float32)
    yy = (np.arange(H, dtype=np.float32)[:, None] + 0.5) / H
    xx = (np.arange(W, dtype=np.float32)[None, :] + 0.5) / W
    fall = np.array([1.0, 0.75, 0.55, 0.4, 0.3, 0.22, 0.16, 0.1], np.float32)
    disp = np.zeros((L, B, H, W, 1), np.float32)
    for b in range(B):
        if cam == 'identity':      # two fronto-parallel planes: 1/3.5 on the left half, 1/2 on the right
            base = np.where(xx < 0.5, 1.0 / 3.5, 0.5) * np.ones((H, 1), np.float32)
        else:                      # road-plane-like ramp plus smooth bumps
            base = max_disp * np.clip((yy - 0.45) / 0.55, 0.02, 1.0) * np.ones((1, W), np.float32)
            base = base + 0.05 * max_disp * np.sin(6.0 * xx + rs.uniform(0, 6.28)) * np.sin(4.0 * yy + rs.uniform(0, 6.28))
            base = np.clip(base, 0.02 * max_disp, 0.98 * max_disp)
        for l in range(L):
            disp[l, b, :, :, 0] = base * fall[l]
    mask = rs.uniform(0.2, 1.0, (L, B, H, W, 1)).astype(np.float32) if cam == 'synth' else np.ones((L, B, H, W, 1), np.float32)
    if cam == 'kitti':
        k = np.stack([kitti_k(H, W)] * B)
        rot = np.stack([np.eye(3, dtype=np.float32)] * B)
        t = np.tile(np.array([[-0.5327], [0.0], [0.0]], np.float32), (B, 1, 1))
        k_s, k_t = k, k.copy()
    elif cam == 'synth':
        k_s = np.stack([synth_k(H, W)] * B)
        k_t = k_s.copy()
        rot = np.stack([_rot(*rs.uniform(-0.05, 0.05, 3)) for _ in range(B)])
        t = rs.uniform(-0.1, 0.1, (B, 3, 1)).astype(np.float32)
    else:
        k_s = np.stack([synth_k(H, W)] * B)
        k_t = k_s.copy()
        rot = np.stack([np.eye(3, dtype=np.float32)] * B)
        t = np.zeros((B, 3, 1), np.float32)
    return dict(tex=tex, mask=mask, disp=disp, k_s=k_s, k_t=k_t, rot=rot, t=t)
