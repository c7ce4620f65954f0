"""Minimal deterministic forward splat renderer (test infrastructure).

The reference hands its SoA to an un-vendored CUDA rasteriser (call site GaussianView.cpp:1099-1128).  To state parity
in image space (north_star: >= 50 dB PSNR on rendered views) the tests render the product's and the oracle's deformed
Gaussians with this same renderer: the standard 3DGS forward pass (EWA projection with the 0.3 px low-pass, degree-3
SH colour, front-to-back alpha blending in depth order, alpha clamp 0.99, 1/255 cut-off), written in torch, fixed
evaluation order, no atomics.  It is not part of the product path."""
import math

import numpy as np
import torch

C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
      1.445305721320277, -0.5900435899266435]


def orbit_cameras(n_views=4, radius=2.6, height=0.6, size=256, fov_deg=40.0):
    """Look-at cameras on a circle around the origin (stand-in for transforms_test.json, which needs the datasets)."""
    cams = []
    f = 0.5 * size / math.tan(0.5 * math.radians(fov_deg))
    for i in range(n_views):
        a = 2 * math.pi * i / n_views + 0.3
        eye = np.array([radius * math.cos(a), radius * math.sin(a), height], np.float64)
        fwd = -eye / np.linalg.norm(eye)
        right = np.cross(fwd, [0, 0, 1.0]); right /= np.linalg.norm(right)
        down = np.cross(fwd, right)
        R = np.stack([right, down, fwd])            # world -> camera (x right, y down, z forward)
        cams.append(dict(R=R, t=-R @ eye, eye=eye, f=f, size=size))
    return cams


def transforms_cameras(path, stride=1):
    """Cameras of a dataset's transforms_test.json as the reference loads them (src/core/scene/ParseData.cpp:243-246,
    src/core/assets/InputCamera.cpp:1513-1576): `transform_matrix` is camera-to-world in the Blender convention (x right, y up,
    -z forward); focal = 0.5 w / tan(camera_angle_x / 2) (= fl_x), w x h pixels.  Returns every `stride`-th frame."""
    import json
    d = json.load(open(path))
    w = int(d.get("w", 800))
    f = 0.5 * w / math.tan(0.5 * float(d["camera_angle_x"]))
    cams = []
    for fr in d["frames"][::stride]:
        c2w = np.array(fr["transform_matrix"], np.float64)
        right, up, back, eye = c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3]
        R = np.stack([right, -up, -back])           # world -> camera (x right, y down, z forward)
        cams.append(dict(R=R, t=-R @ eye, eye=eye, f=f, size=w))
    return cams


def _sh_color(shs, dirs):
    sh = shs.reshape(-1, 16, 3)
    x, y, z = dirs[:, 0:1], dirs[:, 1:2], dirs[:, 2:3]
    xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
    c = C0 * sh[:, 0] - C1 * y * sh[:, 1] + C1 * z * sh[:, 2] - C1 * x * sh[:, 3]
    c = c + C2[0] * xy * sh[:, 4] + C2[1] * yz * sh[:, 5] + C2[2] * (2 * zz - xx - yy) * sh[:, 6] + C2[3] * xz * sh[:, 7] + C2[4] * (xx - yy) * sh[:, 8]
    c = c + C3[0] * y * (3 * xx - yy) * sh[:, 9] + C3[1] * xy * z * sh[:, 10] + C3[2] * y * (4 * zz - xx - yy) * sh[:, 11] \
        + C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[:, 12] + C3[4] * x * (4 * zz - xx - yy) * sh[:, 13] \
        + C3[5] * z * (xx - yy) * sh[:, 14] + C3[6] * x * (xx - 3 * yy) * sh[:, 15]
    return torch.clamp(c + 0.5, min=0.0)


@torch.no_grad()
def render(g, cam, device="cuda", chunk=256):
    """g: dict of numpy arrays pos (N,3), rot (N,4 w,x,y,z), scale (N,3), opacity (N), shs (N,48).  Returns (H, W, 3) float64."""
    dt = torch.float64
    pos = torch.as_tensor(g["pos"], dtype=dt, device=device)
    q = torch.as_tensor(g["rot"], dtype=dt, device=device)
    s = torch.as_tensor(g["scale"], dtype=dt, device=device)
    op = torch.as_tensor(g["opacity"], dtype=dt, device=device)
    shs = torch.as_tensor(g["shs"], dtype=dt, device=device)
    R = torch.as_tensor(cam["R"], dtype=dt, device=device); t = torch.as_tensor(cam["t"], dtype=dt, device=device)
    eye = torch.as_tensor(cam["eye"], dtype=dt, device=device)
    f, size = cam["f"], cam["size"]
    pc = pos @ R.T + t
    keep = pc[:, 2] > 0.2
    pc, q, s, op, shs, pos = pc[keep], q[keep], s[keep], op[keep], shs[keep], pos[keep]
    q = q / q.norm(dim=1, keepdim=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    Rg = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                      2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                      2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], dim=1).reshape(-1, 3, 3)
    M = Rg * s[:, None, :]
    cov3 = M @ M.transpose(1, 2)
    zc = pc[:, 2]
    J = torch.zeros((len(pc), 2, 3), dtype=dt, device=device)
    J[:, 0, 0] = f / zc; J[:, 0, 2] = -f * pc[:, 0] / (zc * zc)
    J[:, 1, 1] = f / zc; J[:, 1, 2] = -f * pc[:, 1] / (zc * zc)
    T = J @ R
    cov2 = T @ cov3 @ T.transpose(1, 2)
    a, b, c = cov2[:, 0, 0] + 0.3, cov2[:, 0, 1], cov2[:, 1, 1] + 0.3
    det = a * c - b * b
    ca, cb, cc = c / det, -b / det, a / det
    mx = f * pc[:, 0] / zc + 0.5 * size
    my = f * pc[:, 1] / zc + 0.5 * size
    dirs = pos - eye
    col = _sh_color(shs, dirs / dirs.norm(dim=1, keepdim=True))
    order = torch.sort(zc, stable=True).indices
    ys, xs = torch.meshgrid(torch.arange(size, dtype=dt, device=device) + 0.5, torch.arange(size, dtype=dt, device=device) + 0.5, indexing="ij")
    img = torch.zeros((size, size, 3), dtype=dt, device=device)
    Tr = torch.ones((size, size), dtype=dt, device=device)
    for i0 in range(0, len(order), chunk):
        idx = order[i0:i0 + chunk]
        dx = xs[..., None] - mx[idx]; dy = ys[..., None] - my[idx]
        power = -0.5 * (ca[idx] * dx * dx + cc[idx] * dy * dy) - cb[idx] * dx * dy
        alpha = torch.clamp(op[idx] * torch.exp(power), max=0.99)
        alpha = torch.where((power > 0) | (alpha < 1.0 / 255.0), torch.zeros_like(alpha), alpha)
        one_m = 1.0 - alpha
        Tbefore = Tr[..., None] * torch.cat([torch.ones_like(one_m[..., :1]), torch.cumprod(one_m, dim=-1)[..., :-1]], dim=-1)
        img += torch.einsum("hwb,bc->hwc", Tbefore * alpha, col[idx])
        Tr = Tr * torch.prod(one_m, dim=-1)
    return img.cpu().numpy()


def psnr(a, b):
    mse = float(np.mean((a - b) ** 2))
    return float("inf") if mse == 0 else 10.0 * math.log10(1.0 / mse)
