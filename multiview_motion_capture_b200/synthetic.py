"""Synthetic multi-camera scenes of the shape BASELINE.json names.

Cameras on a ring looking at the origin, people on a jittered floor grid animated by a smooth random
walk of the BASIC_18 Euler angles, projected to OpenPose BODY_25 detections with pixel noise, joint
dropout, per-view person misses and per-view shuffled person order. The same packed arrays feed the CUDA
path, the oracle and (through oracle/make_golden.py) the real reference.

Packed clip (dict of NumPy arrays):
    kps25  [F, C, Pmax, 25, 3]  float64  BODY_25 (x, y, score), zero padded
    n_pose [F, C]               int32
    K [C,3,3], RT [C,3,4], img_wh [C,2]
    gt_person [F, C, Pmax] int32 (person index behind each detection, -1 = padding)
"""
import numpy as np

B18_PARENTS = np.array([-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 9, 10, 8, 12, 13, 8, 15, 15])
B18_OFFSETS = np.array([
    [0, 0, 0], [0.15, 0, 0], [0, 0, -0.5], [0, 0, -0.5], [-0.15, 0, 0], [0, 0, -0.5], [0, 0, -0.5],
    [0, 0, 0.3], [0, 0, 0.3], [0.2, 0, 0], [0.3, 0, 0], [0.3, 0, 0], [-0.2, 0, 0], [-0.3, 0, 0],
    [-0.3, 0, 0], [0, -0.02, 0.15], [0.07, 0.02, 0.1], [-0.07, 0.02, 0.1]], dtype=np.float64)
LEAF_JOINTS = np.array([3, 6, 11, 14, 16, 17])
# BODY_25 slot <- BASIC_18 joint (others are synthesised or left at score 0)
_B25_FROM_B18 = {0: 15, 1: 8, 2: 12, 3: 13, 4: 14, 5: 9, 6: 10, 7: 11, 8: 0, 9: 4, 10: 5, 11: 6, 12: 1, 13: 2, 14: 3,
                 17: 17, 18: 16}
BODY25_TO_COCO = np.array([0, 16, 15, 18, 17, 5, 2, 6, 3, 7, 4, 12, 9, 13, 10, 14, 11])

GOLDEN_SCENES = {
    # name -> make_clip kwargs (kept tiny: the real reference needs ~0.3 s per track-frame)
    "c4p3": dict(n_views=4, n_people=3, n_frames=9, seed=11, max_poses=4),
    "c8p6": dict(n_views=8, n_people=6, n_frames=5, seed=12, max_poses=8),
    "c8p12": dict(n_views=8, n_people=12, n_frames=3, seed=13, max_poses=12),
}

# tracked (steady-state) goldens at the BASELINE shapes: the reference's tracker is warm-started from the ground truth
# of frame first-1 (oracle/make_golden.py warm), then runs frames first..last. `shelf` = the five Shelf cameras.
WARM_SCENES = {
    "c8p32": dict(kw=dict(n_views=8, n_people=32, n_frames=9, seed=1000, clip_idx=100000), first=3, last=8),
    "c8p16": dict(kw=dict(n_views=8, n_people=16, n_frames=11, seed=1000, clip_idx=200000), first=3, last=10),
    "c8p12": dict(kw=dict(n_views=8, n_people=12, n_frames=10, seed=13, max_poses=12), first=3, last=9),
    "c5p4": dict(kw=dict(n_views=5, n_people=4, n_frames=14, seed=1000, clip_idx=300000), first=3, last=13, shelf=True),
}


def make_warm_scene(name, shelf_calib=None):
    spec = WARM_SCENES[name]
    kw = dict(spec["kw"])
    if spec.get("shelf"):
        if shelf_calib is None:
            import os
            g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                                     "shelf_inputs.npz"))
            shelf_calib = (g["K"], g["RT"], g["img_wh"][0])
        kw["shelf_calib"] = shelf_calib
    return make_clip(**kw)


def body25_to_coco(kps25):
    """OpenPose BODY_25 -> COCO-17 joint gather (reference: src/pose_def.py:262-270)."""
    return kps25[..., BODY25_TO_COCO, :]


def _rot_xyz(e):
    """(...,3) Euler angles -> (...,3,3) with R = Rx(a) Ry(b) Rz(c)."""
    a, b, c = e[..., 0], e[..., 1], e[..., 2]
    ca, sa, cb, sb, cc, sc = np.cos(a), np.sin(a), np.cos(b), np.sin(b), np.cos(c), np.sin(c)
    z, o = np.zeros_like(a), np.ones_like(a)
    rx = np.stack([o, z, z, z, ca, -sa, z, sa, ca], -1).reshape(a.shape + (3, 3))
    ry = np.stack([cb, z, sb, z, o, z, -sb, z, cb], -1).reshape(a.shape + (3, 3))
    rz = np.stack([cc, -sc, z, sc, cc, z, z, z, o], -1).reshape(a.shape + (3, 3))
    return rx @ ry @ rz


def fk_batch(root, euler, scale):
    """root (...,3), euler (...,18,3), scale (...) -> joints (...,18,3) on the reference skeleton."""
    R = _rot_xyz(euler)
    shp = root.shape[:-1]
    pos = np.zeros(shp + (18, 3))
    G = np.zeros(shp + (18, 3, 3))
    G[..., 0, :, :] = R[..., 0, :, :]
    pos[..., 0, :] = root
    for j in range(1, 18):
        p = B18_PARENTS[j]
        G[..., j, :, :] = G[..., p, :, :] @ R[..., j, :, :]
        pos[..., j, :] = pos[..., p, :] + (G[..., p, :, :] @ (B18_OFFSETS[j] * scale[..., None])[..., None])[..., 0]
    return pos


def make_cameras(rng, n_views, img_wh=(1920, 1080), focal=1000.0):
    W, H = img_wh
    K = np.array([[focal, 0, W / 2], [0, focal, H / 2], [0, 0, 1.0]])
    Ks, RTs = [], []
    base = rng.uniform(0, 2 * np.pi)
    for v in range(n_views):
        th = base + 2 * np.pi * v / n_views + rng.uniform(-0.1, 0.1)
        r, h = rng.uniform(5.0, 6.0), rng.uniform(2.4, 3.0)
        c = np.array([r * np.cos(th), r * np.sin(th), h])
        fwd = -c / np.linalg.norm(c)
        right = np.cross(fwd, np.array([0, 0, 1.0]))
        right /= np.linalg.norm(right)
        down = np.cross(fwd, right)
        R = np.stack([right, down, fwd])
        RTs.append(np.concatenate([R, (-R @ c)[:, None]], axis=1))
        Ks.append(K.copy())
    return np.array(Ks), np.array(RTs)


def make_clip(n_views=8, n_people=32, n_frames=10, seed=1000, clip_idx=0, max_poses=None, img_wh=(1920, 1080),
              noise_px=2.0, p_joint_drop=0.05, p_person_miss=0.10, fps=30.0, floor=10.0, shelf_calib=None):
    """One packed clip. `shelf_calib=(K, RT, img_wh)` re-uses given cameras (config 2 uses the Shelf ones)."""
    rng = np.random.default_rng(seed + clip_idx)
    if shelf_calib is not None:
        Ks, RTs, img_wh = np.asarray(shelf_calib[0]), np.asarray(shelf_calib[1]), tuple(shelf_calib[2])
        n_views = len(Ks)
        floor = 4.0
    else:
        Ks, RTs = make_cameras(rng, n_views, img_wh)
    W, H = img_wh
    Pm = max_poses or n_people
    P = np.einsum("vij,vjk->vik", Ks, RTs)
    # people on a jittered grid
    g = int(np.ceil(np.sqrt(n_people)))
    cell = floor / g
    cells = rng.permutation(g * g)[:n_people]
    jit = max(0.0, (cell - 0.8) / 2)
    xy = np.stack([(cells % g + 0.5) * cell - floor / 2, (cells // g + 0.5) * cell - floor / 2], 1)
    xy += rng.uniform(-jit, jit, size=xy.shape) if jit > 0 else 0.0
    root = np.concatenate([xy, np.full((n_people, 1), 0.95)], 1)
    scale = rng.uniform(0.9, 1.1, size=n_people)
    root[:, 2] *= scale
    euler = rng.normal(0, 0.15, size=(n_people, 18, 3)).clip(-0.8, 0.8)
    euler[:, 0, 2] = rng.uniform(-0.8, 0.8, size=n_people)
    euler[:, LEAF_JOINTS] = 0.0
    vel = rng.normal(0, 0.5, size=(n_people, 2))
    kps25 = np.zeros((n_frames, n_views, Pm, 25, 3))
    n_pose = np.zeros((n_frames, n_views), dtype=np.int32)
    gt_person = np.full((n_frames, n_views, Pm), -1, dtype=np.int32)
    gt_joints = np.zeros((n_frames, n_people, 18, 3))
    gt_root = np.zeros((n_frames, n_people, 3))
    gt_euler = np.zeros((n_frames, n_people, 18, 3))
    for f in range(n_frames):
        if f > 0:
            euler = (euler + rng.normal(0, 0.02, size=euler.shape)).clip(-0.8, 0.8)
            euler[:, LEAF_JOINTS] = 0.0
            vel = (vel + rng.normal(0, 0.05, size=vel.shape))
            sp = np.linalg.norm(vel, axis=1, keepdims=True)
            vel = np.where(sp > 1.5, vel * 1.5 / np.maximum(sp, 1e-9), vel)
            root[:, :2] = (root[:, :2] + vel / fps).clip(-floor / 2, floor / 2)
        J = fk_batch(root, euler, scale)  # (N,18,3)
        gt_joints[f] = J
        gt_root[f] = root
        gt_euler[f] = euler
        # 3D points behind the 25 BODY_25 slots
        X = np.zeros((n_people, 25, 3))
        has = np.zeros(25, dtype=bool)
        for b25, b18 in _B25_FROM_B18.items():
            X[:, b25] = J[:, b18]
            has[b25] = True
        ear_axis = J[:, 16] - J[:, 17]
        ear_axis /= np.maximum(np.linalg.norm(ear_axis, axis=1, keepdims=True), 1e-9)
        X[:, 16] = J[:, 15] + 0.03 * ear_axis  # L eye
        X[:, 15] = J[:, 15] - 0.03 * ear_axis  # R eye
        has[[15, 16]] = True
        Xh = np.concatenate([X, np.ones((n_people, 25, 1))], -1)
        for v in range(n_views):
            uvw = Xh @ P[v].T
            z = uvw[..., 2]
            uv = uvw[..., :2] / np.where(np.abs(z) < 1e-9, 1e-9, z)[..., None]
            uv = uv + rng.normal(0, noise_px, size=uv.shape)
            score = rng.uniform(0.6, 0.95, size=(n_people, 25))
            ok = has[None, :] & (z > 0.3) & (uv[..., 0] >= 0) & (uv[..., 0] < W) & (uv[..., 1] >= 0) & (uv[..., 1] < H)
            ok &= rng.uniform(size=ok.shape) >= p_joint_drop
            det = np.concatenate([uv, score[..., None]], -1) * ok[..., None]
            seen = (ok[:, BODY25_TO_COCO].sum(1) >= 6) & (rng.uniform(size=n_people) >= p_person_miss)
            ids = np.nonzero(seen)[0]
            ids = rng.permutation(ids)[:Pm]
            n_pose[f, v] = len(ids)
            kps25[f, v, :len(ids)] = det[ids]
            gt_person[f, v, :len(ids)] = ids
    return dict(kps25=kps25, n_pose=n_pose, K=Ks, RT=RTs, img_wh=np.array([img_wh] * n_views, dtype=np.int32),
                gt_person=gt_person, gt_joints=gt_joints, gt_root=gt_root, gt_euler=gt_euler, gt_scale=scale)


def make_batch(n_clips, n_views, n_people, n_frames, seed=1000, max_poses=None, **kw):
    """B clips stacked for the clip-batch pipeline: kps [F,B,C,Pmax,17,3] COCO, n_pose [F,B,C], K [B,C,3,3],
    RT [B,C,3,4]."""
    clips = [make_clip(n_views, n_people, n_frames, seed, clip_idx=i, max_poses=max_poses, **kw) for i in range(n_clips)]
    kps = np.stack([body25_to_coco(c["kps25"]) for c in clips], 1)
    n_pose = np.stack([c["n_pose"] for c in clips], 1)
    return dict(kps=np.ascontiguousarray(kps), n_pose=np.ascontiguousarray(n_pose), K=np.stack([c["K"] for c in clips]),
                RT=np.stack([c["RT"] for c in clips]), clips=clips)


class SceneStream:
    """The same scene model as make_clip, vectorised over B independent clips and generated frame by frame (bench.py: 1184
    distinct 8 x 32 clips per GPU, or 4096 clips x 600 frames streamed without holding them in memory). Deterministic in
    (seed, n_clips); clip b of a stream is NOT clip b of make_clip (different use of the random stream), the
    distribution is the same. `clip_offset` lets a rank generate its own slice of a larger job's clips: clip b of the
    stream draws from default_rng([seed, clip_offset + b]), so a clip's data does not depend on how clips are sharded."""

    def __init__(self, n_clips, n_views=8, n_people=32, seed=1000, clip_offset=0, img_wh=(1920, 1080), noise_px=2.0,
                 p_joint_drop=0.05, p_person_miss=0.10, fps=30.0, floor=10.0, shelf_calib=None, clip_ids=None):
        B, N = n_clips, n_people
        self.B, self.N, self.fps, self.noise_px, self.p_drop, self.p_miss = B, N, fps, noise_px, p_joint_drop, p_person_miss
        ids = np.asarray(clip_ids) if clip_ids is not None else clip_offset + np.arange(B)
        self.rngs = [np.random.default_rng([seed, int(i)]) for i in ids]
        if shelf_calib is not None:
            K1, RT1, img_wh = np.asarray(shelf_calib[0]), np.asarray(shelf_calib[1]), tuple(int(x) for x in shelf_calib[2])
            self.K, self.RT = np.repeat(K1[None], B, 0), np.repeat(RT1[None], B, 0)
            floor = 4.0
        else:
            cams = [make_cameras(r, n_views, img_wh) for r in self.rngs]
            self.K, self.RT = np.stack([c[0] for c in cams]), np.stack([c[1] for c in cams])
        self.C = self.K.shape[1]
        self.W, self.H = img_wh
        self.floor = floor
        self.P = np.einsum("bvij,bvjk->bvik", self.K, self.RT)
        g = int(np.ceil(np.sqrt(N)))
        cell = floor / g
        jit = max(0.0, (cell - 0.8) / 2)
        cells = np.stack([r.permutation(g * g)[:N] for r in self.rngs])
        xy = np.stack([(cells % g + 0.5) * cell - floor / 2, (cells // g + 0.5) * cell - floor / 2], -1)
        if jit > 0:
            xy = xy + self._draw(lambda r: r.uniform(-jit, jit, size=(N, 2)))
        self.scale = self._draw(lambda r: r.uniform(0.9, 1.1, size=N))
        self.root = np.concatenate([xy, 0.95 * self.scale[..., None]], -1)
        self.euler = self._draw(lambda r: r.normal(0, 0.15, size=(N, 18, 3))).clip(-0.8, 0.8)
        self.euler[:, :, 0, 2] = self._draw(lambda r: r.uniform(-0.8, 0.8, size=N))
        self.euler[:, :, LEAF_JOINTS] = 0.0
        self.vel = self._draw(lambda r: r.normal(0, 0.5, size=(N, 2)))
        self.frame = -1

    def _draw(self, f):
        return np.stack([f(r) for r in self.rngs])

    def next(self):
        """Advance one frame. Returns dict(kps25 [B,C,N,25,3], n_pose [B,C], gt_person [B,C,N], gt_root, gt_euler)."""
        B, N, C, D = self.B, self.N, self.C, self._draw      # (every draw comes from the clip's own generator)
        self.frame += 1
        if self.frame > 0:
            self.euler = (self.euler + D(lambda r: r.normal(0, 0.02, size=(N, 18, 3)))).clip(-0.8, 0.8)
            self.euler[:, :, LEAF_JOINTS] = 0.0
            self.vel = self.vel + D(lambda r: r.normal(0, 0.05, size=(N, 2)))
            sp = np.linalg.norm(self.vel, axis=-1, keepdims=True)
            self.vel = np.where(sp > 1.5, self.vel * 1.5 / np.maximum(sp, 1e-9), self.vel)
            self.root[..., :2] = (self.root[..., :2] + self.vel / self.fps).clip(-self.floor / 2, self.floor / 2)
        J = fk_batch(self.root, self.euler, self.scale)                 # [B,N,18,3]
        X = np.zeros((B, N, 25, 3))
        has = np.zeros(25, dtype=bool)
        for b25, b18 in _B25_FROM_B18.items():
            X[:, :, b25] = J[:, :, b18]
            has[b25] = True
        ear = J[:, :, 16] - J[:, :, 17]
        ear /= np.maximum(np.linalg.norm(ear, axis=-1, keepdims=True), 1e-9)
        X[:, :, 16] = J[:, :, 15] + 0.03 * ear
        X[:, :, 15] = J[:, :, 15] - 0.03 * ear
        has[[15, 16]] = True
        Xh = np.concatenate([X, np.ones((B, N, 25, 1))], -1)
        uvw = np.einsum("bnjk,bvik->bvnji", Xh, self.P)                 # [B,C,N,25,3]
        z = uvw[..., 2]
        uv = uvw[..., :2] / np.where(np.abs(z) < 1e-9, 1e-9, z)[..., None]
        uv = uv + D(lambda r: r.normal(0, self.noise_px, size=(C, N, 25, 2)))
        score = D(lambda r: r.uniform(0.6, 0.95, size=(C, N, 25)))
        ok = has & (z > 0.3) & (uv[..., 0] >= 0) & (uv[..., 0] < self.W) & (uv[..., 1] >= 0) & (uv[..., 1] < self.H)
        ok &= D(lambda r: r.uniform(size=(C, N, 25))) >= self.p_drop
        det = np.concatenate([uv, score[..., None]], -1) * ok[..., None]
        seen = (ok[..., BODY25_TO_COCO].sum(-1) >= 6) & (D(lambda r: r.uniform(size=(C, N))) >= self.p_miss)
        key = np.where(seen, D(lambda r: r.uniform(size=(C, N))), 2.0)          # random order of the seen people, unseen last
        order = np.argsort(key, axis=-1)
        n_pose = seen.sum(-1).astype(np.int32)
        kps25 = np.take_along_axis(det, order[..., None, None], axis=2)
        live = np.arange(N)[None, None, :] < n_pose[..., None]
        kps25 = kps25 * live[..., None, None]
        gt_person = np.where(live, order, -1).astype(np.int32)
        return dict(kps25=kps25, n_pose=n_pose, gt_person=gt_person, gt_root=self.root.copy(), gt_euler=self.euler.copy(),
                    gt_joints=J)

    def gt_params(self, side_bone_lens):
        """[B,N,68] pose parameters of the CURRENT frame's ground truth (root, euler, scaled side bone lengths)."""
        lens = np.asarray(side_bone_lens)[None, None, :] * self.scale[..., None]
        return np.concatenate([self.root, self.euler.reshape(self.B, self.N, 54), lens], -1)


class DeviceSceneStream:
    """SceneStream's scene model with the per-frame work done by torch on the device the clips are tracked on, so that a
    long job (BASELINE config 5: 4096 clips x 600 frames) can synthesise its detections without the host: every random
    number is a counter-based hash (splitmix64) of (seed, clip id, frame, draw, element), so a clip's data depends on
    nothing but its id - not on the batch it is generated in, nor on the rank. The scene itself (cameras, people, start
    poses) is SceneStream's. Benchmark input synthesis only; not part of the capture path."""

    _M1, _M2, _G = -4658895280553007687, -7723592293110705685, -7046029254386353131   # splitmix64 constants as int64

    def __init__(self, clip_ids, n_views=8, n_people=32, seed=1000, device="cuda", shuffle=True, **kw):
        import torch
        self.t = torch
        self.shuffle = shuffle     # False: pose slot p of every view is person p (an unseen person is an all-zero pose)
        host = SceneStream(len(clip_ids), n_views, n_people, seed=seed, clip_ids=clip_ids, **kw)
        self.B, self.N, self.C = host.B, host.N, host.C
        self.W, self.H, self.floor, self.fps = host.W, host.H, host.floor, host.fps
        self.noise_px, self.p_drop, self.p_miss = host.noise_px, host.p_drop, host.p_miss
        self.K, self.RT = host.K, host.RT
        d = lambda a: torch.as_tensor(a, dtype=torch.float64, device=device)
        self.P, self.root, self.euler, self.scale, self.vel = d(host.P), d(host.root), d(host.euler), d(host.scale), d(host.vel)
        self.ids = torch.as_tensor(np.asarray(clip_ids, dtype=np.int64), device=device)
        self.seed, self.device, self.frame = int(seed), device, -1
        self.off = d(B18_OFFSETS)
        self.leaf = torch.as_tensor(LEAF_JOINTS, device=device)
        self.b25_src = torch.as_tensor([_B25_FROM_B18.get(j, 0) for j in range(25)], device=device)
        self.has = torch.zeros(25, dtype=torch.bool, device=device)
        self.has[list(_B25_FROM_B18) + [15, 16]] = True
        self.coco = torch.as_tensor(BODY25_TO_COCO, device=device)

    def _mix(self, z):
        t = self.t
        z = z + self._G
        z = (z ^ ((z >> 30) & ((1 << 34) - 1))) * self._M1
        z = (z ^ ((z >> 27) & ((1 << 37) - 1))) * self._M2
        return z ^ ((z >> 31) & ((1 << 33) - 1))

    def _uniform(self, draw, shape):
        """[B, *shape] uniforms in [0, 1): hash of (seed, clip id, frame, draw, element index)."""
        t = self.t
        n = int(np.prod(shape))
        key = self._mix(self.ids * 1000003 + self.seed) ^ self._mix(t.full_like(self.ids, self.frame * 64 + draw))
        z = self._mix(key[:, None] * 2654435761 + t.arange(n, device=self.device, dtype=t.int64)[None, :])
        u = ((z >> 11) & ((1 << 53) - 1)).to(t.float64) * (1.0 / (1 << 53))
        return u.reshape((self.B,) + tuple(shape))

    def _normal(self, draw, shape, sigma):
        t = self.t
        u1, u2 = self._uniform(draw, shape), self._uniform(draw + 32, shape)
        return sigma * t.sqrt(-2.0 * t.log(1.0 - u1)) * t.cos(2.0 * np.pi * u2)

    def _fk(self):
        t = self.t
        a, b, c = self.euler[..., 0], self.euler[..., 1], self.euler[..., 2]
        ca, sa, cb, sb, cc, sc = t.cos(a), t.sin(a), t.cos(b), t.sin(b), t.cos(c), t.sin(c)
        z, o = t.zeros_like(a), t.ones_like(a)
        rx = t.stack([o, z, z, z, ca, -sa, z, sa, ca], -1).reshape(a.shape + (3, 3))
        ry = t.stack([cb, z, sb, z, o, z, -sb, z, cb], -1).reshape(a.shape + (3, 3))
        rz = t.stack([cc, -sc, z, sc, cc, z, z, z, o], -1).reshape(a.shape + (3, 3))
        R = rx @ ry @ rz                                           # [B,N,18,3,3]
        pos, G = [self.root], [R[:, :, 0]]
        for j in range(1, 18):
            p = int(B18_PARENTS[j])
            off = self.off[j][None, None, :] * self.scale[..., None]
            pos.append(pos[p] + (G[p] @ off[..., None])[..., 0])
            G.append(G[p] @ R[:, :, j])
        return t.stack(pos, 2)                                     # [B,N,18,3]

    def next(self):
        """Advance one frame: (kps [B,C,N,17,3] float64 COCO detections, n_pose [B,C] int32) on the device."""
        t = self.t
        B, N, C = self.B, self.N, self.C
        self.frame += 1
        if self.frame > 0:
            self.euler = (self.euler + self._normal(0, (N, 18, 3), 0.02)).clamp(-0.8, 0.8)
            self.euler[:, :, self.leaf] = 0.0
            self.vel = self.vel + self._normal(1, (N, 2), 0.05)
            sp = self.vel.norm(dim=-1, keepdim=True)
            self.vel = t.where(sp > 1.5, self.vel * 1.5 / sp.clamp_min(1e-9), self.vel)
            self.root[..., :2] = (self.root[..., :2] + self.vel / self.fps).clamp(-self.floor / 2, self.floor / 2)
        J = self._fk()
        X = J[:, :, self.b25_src]                                  # [B,N,25,3]
        ear = J[:, :, 16] - J[:, :, 17]
        ear = ear / ear.norm(dim=-1, keepdim=True).clamp_min(1e-9)
        X[:, :, 16] = J[:, :, 15] + 0.03 * ear
        X[:, :, 15] = J[:, :, 15] - 0.03 * ear
        Xh = t.cat([X, t.ones((B, N, 25, 1), dtype=t.float64, device=self.device)], -1)
        uvw = t.einsum("bnjk,bvik->bvnji", Xh, self.P)             # [B,C,N,25,3]
        z = uvw[..., 2]
        uv = uvw[..., :2] / t.where(z.abs() < 1e-9, t.full_like(z, 1e-9), z)[..., None]
        uv = uv + self._normal(2, (C, N, 25, 2), self.noise_px)
        score = 0.6 + 0.35 * self._uniform(3, (C, N, 25))
        ok = self.has & (z > 0.3) & (uv[..., 0] >= 0) & (uv[..., 0] < self.W) & (uv[..., 1] >= 0) & (uv[..., 1] < self.H)
        ok = ok & (self._uniform(4, (C, N, 25)) >= self.p_drop)
        det = t.cat([uv, score[..., None]], -1) * ok[..., None]
        seen = (ok[..., self.coco].sum(-1) >= 6) & (self._uniform(5, (C, N)) >= self.p_miss)
        key = t.where(seen, self._uniform(6, (C, N)), t.full((B, C, N), 2.0, dtype=t.float64, device=self.device))
        order = key.argsort(dim=-1)
        n_pose = seen.sum(-1).to(t.int32)
        det = det[..., self.coco, :]                               # COCO-17
        if not self.shuffle:
            return (det * seen[..., None, None]).contiguous(), t.full((B, C), N, dtype=t.int32, device=self.device)
        kps = t.take_along_dim(det, order[..., None, None], dim=2)
        live = t.arange(N, device=self.device)[None, None, :] < n_pose[..., None]
        return (kps * live[..., None, None]).contiguous(), n_pose.contiguous()
