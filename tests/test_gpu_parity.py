"""GPU parity: the CUDA path (through the C ABI, libpbr_b200.so) against the CPU oracle, bit for bit.

Coverage, depth ordering and colours are integer / deterministic-fp32 work, so the bar is exact
equality of every output byte on the same inputs; the golden notebook tiles are compared directly
as well (tolerance: the documented 1-LSB .5-boundary pixels of tile 1).
"""
import numpy as np
import pytest
import torch

from util import (cartpole_states, config5_renderer, host_threads, many_cubes_renderer, mixed_mesh_renderer,
                  oracle_render, oracle_subset)

pytestmark = pytest.mark.gpu


def _assert_same(gpu: torch.Tensor, ref: np.ndarray, what: str = ""):
    got = gpu.cpu().numpy()
    assert got.shape == ref.shape and got.dtype == np.uint8
    if not np.array_equal(got, ref):
        diff = (got != ref).any(1)
        bad_scenes = np.argwhere(diff.any((1, 2))).ravel()
        raise AssertionError(f"{what}: {int(diff.sum())} pixels differ in scenes {bad_scenes[:10].tolist()} "
                             f"(first at {np.argwhere(diff)[0].tolist()})")


def _cartpole(n, tile=(64, 64), general=False, **kw):
    from pybatchrender_b200.envs.cartpole import CartPoleRenderer
    r = CartPoleRenderer(dict(num_scenes=n, tile_resolution=tile, device="cuda", **kw))
    if general:
        r.render_flags = 1          # PBR_FRAME_FORCE_GENERAL: skip the one-warp-per-scene kernel
    return r


def test_native_library_is_loaded():
    import pybatchrender_b200._native as nat
    assert nat.load().pbr_version() >= 100
    with open("/proc/self/maps") as f:
        assert "libpbr_b200.so" in f.read()


def test_config1_cartpole_4_scenes_vs_oracle_and_golden_states(golden):
    r = _cartpole(4)
    state = torch.zeros(4, 4)
    state[:3] = torch.tensor(golden["obs3"])
    px = r.step(state.cuda())
    assert px.shape == (4, 3, 64, 64) and px.dtype == torch.uint8 and px.is_cuda and px.is_contiguous()
    _assert_same(px, oracle_render(r), "config 1")
    # non-trivial image: rail + cart + pole visible
    assert (px[0] != 0).any()


def test_notebook_golden_tiles_on_gpu(golden):
    """Same scene as the notebook (N=4098 -> aspect 65/64, bg 105): tiles 0 and 2 exact, tile 1 <= 1 LSB."""
    r = _cartpole(4098)
    r.set_background_color(0.41, 0.41, 0.41)
    state = torch.zeros(4098, 4)
    state[:3] = torch.tensor(golden["obs3"])
    px = r.step(state.cuda()).cpu().numpy()
    gold = golden["initial"].transpose(0, 3, 1, 2)
    assert np.array_equal(px[0], gold[0])
    assert np.array_equal(px[2], gold[2])
    d = np.abs(px[1].astype(int) - gold[1].astype(int))
    assert d.max() <= 1 and (d.sum(0) > 0).sum() <= 25
    assert np.array_equal((px[1] != 105).any(0), (gold[1] != 105).any(0))


@pytest.mark.parametrize("general", [False, True], ids=["warp", "general"])
@pytest.mark.parametrize("n,tile", [(256, (64, 64)), (100, (84, 84)), (37, (128, 128)), (9, (50, 30)),
                                    (5, (256, 256)), (3, (8, 8)), (2, (200, 40)), (33, (96, 64))])
def test_cartpole_random_states_bit_exact(n, tile, general):
    r = _cartpole(n, tile, general=general)
    px = r.step(cartpole_states(n, seed=n).cuda())
    _assert_same(px, oracle_render(r), f"cartpole {n}x{tile}")


def test_pose_in_kernel_equals_pose_kernel_equals_generic_setters():
    """CartPole binds cart and pole to the state columns (``PBRNode.set_pose``) and the small-scene kernel
    computes their matrices itself.  Three ways to the same matrices: (a) written by the raster kernel
    (PBR_FRAME_WRITE_MATS), (b) materialised by ``pbr_compose_transforms`` when ``matbuf`` is read, (c) the
    reference's sequence of generic setters (torch ops + ``pbr_pack_transforms``).  (a) == (b) bit for bit --
    the oracle is fed (b), so this is what makes every pixel test also a test of the in-kernel pose;
    (c) agrees within float rounding of sin / cos."""
    n = 512
    r = _cartpole(n)
    st = cartpole_states(n, seed=3).cuda()
    r._step(st)
    assert r.cart._pose is not None and r.pole._pose is not None
    for node in (r.cart, r.pole):
        node._matbuf.fill_(float("nan"))
    r.render(flags=4)                                   # PBR_FRAME_WRITE_MATS
    in_kernel = r.cart._matbuf.clone(), r.pole._matbuf.clone()
    assert not torch.isnan(in_kernel[0]).any() and not torch.isnan(in_kernel[1]).any()
    composed = r.cart.matbuf.clone(), r.pole.matbuf.clone()
    assert torch.equal(in_kernel[0], composed[0]) and torch.equal(in_kernel[1], composed[1])
    # without the flag the small-scene kernel leaves the buffers alone
    for node in (r.cart, r.pole):
        node._matbuf.fill_(7.0)
    r.render()
    assert (r.cart._matbuf == 7.0).all() and (r.pole._matbuf == 7.0).all()
    # mirrors follow the pose (reference node.py:116-134 keeps transforms_b44 current on every setter)
    assert torch.equal(r.cart.transforms_b44[:, 0, 3], st[:, 0])
    torch.testing.assert_close(r.pole.rot3_b33[:, 0, 2], torch.sin(st[:, 2]), atol=1e-6, rtol=0)
    # (c): the generic path; it ends the binding
    native, r._native = r._native, None
    try:
        r._step(st)
    finally:
        r._native = native
    assert r.cart._pose is None and r.pole._pose is None
    assert torch.equal(r.cart.matbuf, composed[0])                       # translation only: exact
    torch.testing.assert_close(r.pole.matbuf, composed[1], atol=2e-7, rtol=0)
    _assert_same(r.render(), oracle_render(r), "generic setters after a pose")


def test_pose_matrices_equal_the_reference_under_stubs():
    """The reference's own ``PBRNode`` (run under Panda3D stubs, ``tests/golden/make_host_golden.py``) uploaded
    these cart / pole matrices for the notebook's three printed states; the device pose must reproduce them
    (cart: exactly; pole: within float rounding of torch's sin / cos against CUDA's sincosf)."""
    import os
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "host_math.npz"))
    for n in (4, 37, 4098):
        k = gold[f"pole_mat_{n}"].shape[0]
        r = _cartpole(n)
        r._step(torch.tensor(gold[f"state_{n}"]).cuda())
        np.testing.assert_array_equal(r.cart.matbuf[:k].cpu().numpy(), gold[f"cart_mat_{n}"])
        np.testing.assert_allclose(r.pole.matbuf[:k].cpu().numpy(), gold[f"pole_mat_{n}"], atol=2e-7)


def test_more_posed_nodes_than_the_kernel_takes_and_posed_shared_node():
    """Six posed single-triangle-pair nodes (> MAX_FRAME_POSES = 4: the library materialises them with the pose
    kernel first) and a posed shared node (which stays out of the static layer)."""
    from pybatchrender_b200 import PBRRenderer, meshes
    n = 40
    r = PBRRenderer(dict(num_scenes=n, tile_resolution=(64, 64), device="cuda"))
    rng = np.random.default_rng(3)
    nodes = [r.add_node("models/box", instances_per_scene=1, model_pivot_relative_point=(0.5, 0.5, 0.5),
                        shared_across_scenes=(i == 0)) for i in range(3)]
    cam = r.add_camera()
    cam.set_positions(torch.tensor([0.0, -9.0, 1.0]))
    cam.look_at(torch.tensor([0.0, 0.0, 0.0]))
    r.add_light()
    r.setup_environment()
    chans = torch.tensor(rng.uniform(-2.0, 2.0, (n, 8)), dtype=torch.float32).cuda()
    one = torch.tensor(rng.uniform(-2.0, 2.0, (1, 8)), dtype=torch.float32).cuda()
    for i, node in enumerate(nodes):
        c = one if node.shared_across else chans
        node.set_colors(torch.tensor(np.concatenate([rng.uniform(0.2, 1, (node.buf_instances, 3)),
                                                     np.ones((node.buf_instances, 1))], 1), dtype=torch.float32))
        node.set_pose(pos=(c[:, i], c[:, i + 1], 0.3 * i), hpr=(c[:, i + 2], 0.4, c[:, i + 3]), scale=1.0 + 0.25 * i)
    _assert_same(r.render(), oracle_render(r), "three posed nodes (one shared)")
    assert r._base_sig is None          # a posed node follows tensors that change behind the API: never in the static layer
    r.static_layer = False
    _assert_same(r.render(), oracle_render(r), "three posed nodes, no static layer")
    # > MAX_FRAME_POSES posed nodes in the frame
    r2 = PBRRenderer(dict(num_scenes=n, tile_resolution=(64, 64), device="cuda"))
    quad = meshes.box()
    quad.idx = quad.idx[:4].copy()
    meshes.register_mesh("test/two_faces", quad)
    many = [r2.add_node("test/two_faces", instances_per_scene=1) for _ in range(6)]
    cam = r2.add_camera()
    cam.set_positions(torch.tensor([0.0, -9.0, 1.0]))
    r2.add_light()
    r2.setup_environment()
    for i, node in enumerate(many):
        node.set_pose(pos=(chans[:, i], 0.5 * i - 1.0, chans[:, (i + 1) % 8]), hpr=(chans[:, (i + 2) % 8], chans[:, (i + 3) % 8], 0.0),
                      scale=torch.abs(chans[:, (i + 4) % 8]) + 0.5)
    _assert_same(r2.render(), oracle_render(r2), "six posed nodes")


# the three ways a large scene can go: band-based (geometry pre-pass + TMA-staged raster per band of rows), block
# lists (one warp per 8x8 block), geometry fused into the general raster kernel
LARGE_PATHS = {"staged": 8, "binned": 16, "fused": 2}      # PBR_FRAME_FORCE_STAGED / _BINNED / _FUSED


@pytest.mark.parametrize("path", list(LARGE_PATHS))
@pytest.mark.parametrize("channels", [3, 4])
def test_many_cubes_clipping_and_multipass(channels, path):
    """16 boxes/scene = 192 triangle slots (> one 128-slot pass), camera inside the cloud so that
    triangles cross the near plane (clip path) and the guard band."""
    r = many_cubes_renderer(num_scenes=24, instances=16, tile=(64, 64), device="cuda", channels=channels)
    r.render_flags = LARGE_PATHS[path]
    px = r.step()
    ref = oracle_render(r)
    assert (ref != 0).any()
    _assert_same(px, ref, "many cubes")


@pytest.mark.parametrize("path", list(LARGE_PATHS))
@pytest.mark.parametrize("n,inst,tile,seed", [(6, 40, (128, 128), 7), (3, 150, (256, 256), 8), (5, 300, (84, 84), 9),
                                              (2, 64, (200, 120), 10), (3, 90, (75, 53), 11)])
def test_many_cubes_large_tile_bands(n, inst, tile, seed, path):
    """Several bands per tile and several 128-record chunks per scene: geometry pre-pass + TMA-staged
    raster against the fused kernel and the oracle."""
    r = many_cubes_renderer(num_scenes=n, instances=inst, tile=tile, device="cuda", seed=seed)
    r.render_flags = LARGE_PATHS[path]
    _assert_same(r.step(), oracle_render(r), f"many cubes {tile} x{inst} {path}")
    assert r._native.device_status(torch.cuda.current_device()) == 0


@pytest.mark.parametrize("path", list(LARGE_PATHS))
def test_camera_inside_geometry_heavy_clipping(path):
    """Huge near triangles: block boxes that span the whole tile (walked by whole warps when the block lists are
    built), int64 records, fans of clipped triangles."""
    r = many_cubes_renderer(num_scenes=16, instances=12, tile=(64, 64), device="cuda", seed=11, spread=3.0,
                            eye=(0.0, -1.0, 0.0))
    r.render_flags = LARGE_PATHS[path]
    _assert_same(r.step(), oracle_render(r), f"heavy clipping {path}")
    big = many_cubes_renderer(num_scenes=4, instances=20, tile=(200, 136), device="cuda", seed=12, spread=4.0,
                              eye=(0.0, -1.5, 0.2))
    big.render_flags = LARGE_PATHS[path]
    _assert_same(big.step(), oracle_render(big), f"heavy clipping, large tile, {path}")


def test_block_list_overflow_is_reported_and_heals():
    """A frame whose (block, record) lists do not fit the scratch sized for it sets the overflow flag; the next
    large-scene call reports PBR_EOVERFLOW once and from then on this device uses the band-based path with
    worst-case record capacity, which renders the same frame exactly."""
    import pybatchrender_b200._native as nat
    dev = torch.cuda.current_device()
    # every triangle is huge (camera inside big boxes): each record lands in nearly every block list
    r = many_cubes_renderer(num_scenes=2, instances=120, tile=(256, 256), device="cuda", seed=13, spread=0.6,
                            eye=(0.0, 0.0, 0.0), two_sided=True)
    node = r._pbr_nodes[0]
    node.set_scales(torch.full((node.buf_instances, 1), 9.0))
    r.render_flags = 16                                # PBR_FRAME_FORCE_BINNED
    r._native.device_status(dev, clear=True)
    r.render()
    torch.cuda.synchronize()
    if r._native.device_status_nosync(dev) & 2:
        with pytest.raises(nat.NativeError, match="dropped triangles"):
            r.render()
        r.render_flags = 0
        _assert_same(r.render(), oracle_render(r), "after the overflow: band-based path, worst-case capacity")
    else:                                              # lists were large enough after all: the frame must be exact
        _assert_same(r.render(), oracle_render(r), "block lists fit")
    r._native.device_status(dev, clear=True)


@pytest.mark.parametrize("general", [False, True], ids=["warp", "general"])
@pytest.mark.parametrize("seed,eye,spread", [(21, (0.0, -1.0, 0.0), 2.0), (22, (0.0, -2.5, 0.3), 3.0),
                                              (23, (0.2, 0.1, 0.0), 2.5), (24, (0.0, -6.0, 0.0), 6.0)])
def test_small_scene_with_clipping(general, seed, eye, spread):
    """3 boxes = 36 slots: eligible for the one-warp-per-scene kernel, camera close enough that
    triangles cross the near plane (fan triangles go to the spare record slots)."""
    r = many_cubes_renderer(num_scenes=64, instances=3, tile=(64, 64), device="cuda", seed=seed, spread=spread,
                            eye=eye)
    if general:
        r.render_flags = 1
    px = r.step()
    ref = oracle_render(r)
    assert (ref != 0).any()
    _assert_same(px, ref, "small scene clipping")
    assert r._native.device_status(torch.cuda.current_device()) == 0


@pytest.mark.parametrize("tile", [(80, 80), (200, 200)], ids=["tma", "warp-copy"])
def test_small_scene_with_huge_triangles_int64_records(tile):
    """Boxes larger than the tile (not clipped: inside the guard band, in front of the near plane) give
    records whose edge functions do not fit an int over their hull: the int64 path.  The small-scene
    kernel keeps that path out of its pair loop with a per-scene flag in the queue items, so scenes
    with and without such records are mixed inside every CTA here."""
    n = 96
    r = many_cubes_renderer(num_scenes=n, instances=2, tile=tile, device="cuda", seed=31, spread=2.0,
                            eye=(0.0, -14.0, 0.0))
    node = r._pbr_nodes[0]
    rng = np.random.default_rng(77)
    big = rng.random(n) < 0.5
    sc = np.where(np.repeat(big, 2), rng.uniform(5.0, 9.0, 2 * n), rng.uniform(0.4, 1.5, 2 * n))
    node.set_scales(torch.tensor(sc.reshape(-1, 1), dtype=torch.float32))
    px = r.step()
    ref = oracle_render(r)
    covered = (ref != 0).any(1).reshape(n, -1).mean(1)
    assert covered.max() > 0.6 and covered.min() < 0.2          # some scenes filled, some nearly empty
    _assert_same(px, ref, "huge triangles")
    assert r._native.device_status(torch.cuda.current_device()) == 0


def test_shared_node_many_instances():
    r = many_cubes_renderer(num_scenes=10, instances=9, tile=(64, 64), device="cuda", seed=5, shared=True)
    px = r.step()
    _assert_same(px, oracle_render(r), "shared I>1")
    assert torch.equal(px[0], px[9])        # same instances and same camera in every scene


def test_two_sided_mesh():
    r = many_cubes_renderer(num_scenes=8, instances=6, tile=(64, 64), device="cuda", seed=9, spread=5.0,
                            two_sided=True)
    _assert_same(r.step(), oracle_render(r), "two sided")


def test_unlit_when_no_light():
    r = many_cubes_renderer(num_scenes=4, instances=4, tile=(32, 32), device="cuda", seed=2, spread=4.0,
                            light=False)
    _assert_same(r.step(), oracle_render(r), "unlit")


def test_scene_window_and_out_reuse():
    r = _cartpole(32)
    r._step(cartpole_states(32, seed=1).cuda())
    full = r.render()
    out = torch.full_like(full, 77)
    r.render(out=out, scene_begin=8, scene_count=16)
    assert torch.equal(out[8:24], full[8:24])
    assert (out[:8] == 77).all() and (out[24:] == 77).all()


def test_deterministic_and_stream_ordered():
    r = _cartpole(512)
    st = cartpole_states(512, seed=4).cuda()
    a = r.step(st).clone()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        b = r.step(st)
    s.synchronize()
    assert torch.equal(a, b)


@pytest.mark.parametrize("tile", [(64, 64), (32, 32)], ids=["tma", "warp-copy"])
def test_back_to_back_steps_into_one_buffer_stay_ordered(tile):
    """The pose kernel and the small-scene raster kernel are launched as programmatic dependents of
    each other (raster(n) -> pose(n+1) -> raster(n+1): the next kernel's CTAs are scheduled while the
    previous one drains).  Frames written back to back into ONE output buffer must not leak into each
    other: after a burst of steps with different states -- eager and replayed from a CUDA graph -- the
    buffer holds exactly the last state's frame."""
    n = 4096
    r = _cartpole(n, tile=tile)
    states = [cartpole_states(n, seed=40 + i).cuda() for i in range(6)]
    out = torch.empty((n, 3, tile[1], tile[0]), dtype=torch.uint8, device="cuda")
    want = [r.step(s).clone() for s in states]
    torch.cuda.synchronize()
    for rounds in range(3):
        for i, s in enumerate(states):
            r.step(s, out=out)
        assert torch.equal(out, want[-1]), f"eager burst {rounds}"
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        r.step(states[0], out=out)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i, s in enumerate(states):
            r.step(s, out=out)
        for i in (3, 1, 4):
            r.step(states[i], out=out)
    for rounds in range(5):
        g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, want[4]), "graph replay"
    # and each frame of a burst, read back by a copy enqueued right behind it
    snaps = []
    for i, s in enumerate(states):
        snaps.append(r.step(s, out=out).clone())
    torch.cuda.synchronize()
    for i in range(len(states)):
        assert torch.equal(snaps[i], want[i]), f"frame {i} of a burst"


@pytest.mark.parametrize("seed", [0, 1])
def test_config2_every_scene_of_4096(seed):
    """BASELINE config 2 (the bench workload) at full occupancy: all 4096 scenes of 64x64 against the oracle --
    the hazards of the <14, true> kernel (bulk stores against byte patches, per-SM store turns, the late
    stores-complete flag, two CTAs per SM, frames overlapping through the programmatic launch chain) only
    exist when every SM is full."""
    n = 4096
    r = _cartpole(n)
    px = r.step(cartpole_states(n, seed=seed).cuda())
    assert px.shape == (n, 3, 64, 64)
    _assert_same(px, oracle_render(r, n_threads=host_threads()), "config 2, every scene")
    nonbg = (px != 0).any(1).flatten(1).sum(1)
    assert int(nonbg.min()) > 20 and int(nonbg.max()) < 600


def test_graph_replay_over_state_and_output_rings_like_the_bench():
    """What bench.py times: CUDA graphs of 16 steps over a ring of 16 states and a ring of 4 output buffers
    (consecutive frames overlap through the programmatic launch chain and write different buffers; every
    fourth frame reuses a buffer).  After several replays every ring buffer must hold exactly the oracle's
    frame of the last state rendered into it; then the same through an eager loop."""
    n, ring, out_ring = 4096, 16, 4
    r = _cartpole(n)
    states = [cartpole_states(n, seed=100 + i).cuda() for i in range(ring)]
    outs = [torch.zeros((n, 3, 64, 64), dtype=torch.uint8, device="cuda") for _ in range(out_ring)]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(3):
            r.step(states[i], out=outs[i % out_ring])
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(ring):
            r.step(states[i], out=outs[i % out_ring])
    for _ in range(4):
        g.replay()
    torch.cuda.synchronize()
    want = {}
    for b in range(out_ring):
        last = max(i for i in range(ring) if i % out_ring == b)
        r._step(states[last])
        want[b] = oracle_render(r, n_threads=host_threads())
        _assert_same(outs[b], want[b], f"graph replay, ring buffer {b}")
    for o in outs:
        o.zero_()
    for rounds in range(3):
        for i in range(ring):
            r.step(states[i], out=outs[i % out_ring])
    torch.cuda.synchronize()
    for b in range(out_ring):
        _assert_same(outs[b], want[b], f"eager loop, ring buffer {b}")
    # ping-pong between two buffers: frame n+1 writes the buffer of frame n-1
    for o in outs[:2]:
        o.zero_()
    for rounds in range(4):
        for i in range(ring):
            r.step(states[i], out=outs[i % 2])
    torch.cuda.synchronize()
    for b in range(2):
        r._step(states[ring - 2 + b])
        _assert_same(outs[b], oracle_render(r, n_threads=host_threads()), f"two-buffer ping-pong, buffer {b}")


def test_cached_frame_description_tracks_every_change():
    """``render`` reuses the native frame description of the previous frame while nothing but tensor contents
    changed; every setter that changes what the description says must invalidate it."""
    n = 48
    r = _cartpole(n)
    st = [cartpole_states(n, seed=70 + i).cuda() for i in range(3)]
    a = r.step(st[0])
    assert r._call_cache is not None
    _assert_same(a, oracle_render(r), "first frame")
    b = r.step(st[1])                                    # same description, other state tensor, new output
    assert b.data_ptr() != a.data_ptr()
    _assert_same(b, oracle_render(r), "cached description, new state")
    st[1][:, 0] += 0.25                                  # contents of the bound tensor change in place
    _assert_same(r.render(), oracle_render(r), "in-place state update")
    r.cart.set_colors(torch.rand(n, 4).cuda())           # node contents through a setter
    _assert_same(r.render(), oracle_render(r), "colours")
    r._pbr_light.set_ambient((0.4, 0.1, 0.3))
    _assert_same(r.render(), oracle_render(r), "light")
    r.set_background_color(0.1, 0.2, 0.3)
    _assert_same(r.render(), oracle_render(r), "background")
    r._pbr_cam.set_positions(torch.tensor([4.0, 6.0, 2.5]))
    r._pbr_cam.look_at(torch.tensor([0.0, 0.0, 0.2]))
    _assert_same(r.render(), oracle_render(r), "camera")
    r.rail.set_positions(torch.tensor([[0.0, 0.1, -0.05]]))            # shared node: static layer re-rendered
    _assert_same(r.render(), oracle_render(r), "shared node")
    r.pole.set_hprs(torch.zeros(n, 3))                   # a generic setter ends the pose binding of that node
    assert r.pole._pose is None
    _assert_same(r.render(), oracle_render(r), "binding ended")
    r.step(st[2])                                        # ... and CartPole's _step binds it again
    assert r.pole._pose is not None
    _assert_same(r.render(), oracle_render(r), "bound again")
    out = torch.full((n, 3, 64, 64), 9, dtype=torch.uint8, device="cuda")
    r.render(out=out, scene_begin=5, scene_count=7)      # scene window through the cached description
    full = r.render()
    assert torch.equal(out[5:12], full[5:12]) and (out[:5] == 9).all() and (out[12:] == 9).all()
    extra = r.add_node("models/box", instances_per_scene=1, model_scale=(0.3, 0.3, 0.3))     # scene structure
    extra.set_positions(torch.tensor(np.random.default_rng(0).uniform(-1, 1, (n, 3)), dtype=torch.float32))
    _assert_same(r.render(), oracle_render(r), "node added")
    extra.np.removeNode()
    _assert_same(r.render(), oracle_render(r), "node removed")


def test_errors_are_loud():
    import pybatchrender_b200._native as nat
    r = _cartpole(4)
    with pytest.raises(ValueError):
        r.render(out=torch.empty((3, 3, 64, 64), dtype=torch.uint8, device="cuda"))
    with pytest.raises(nat.NativeError):
        r._native.render(num_scenes=4, tile_w=64, tile_h=64, channels=5, vp=r._pbr_cam.viewbuf,
                         nodes=r._native_nodes(), out=torch.empty((4, 5, 64, 64), dtype=torch.uint8, device="cuda"),
                         bg=(0, 0, 0, 1), ambient=(0, 0, 0), dir_dir=(0, 0, 1), dir_col=(1, 1, 1), strength=1.0)
    with pytest.raises(nat.NativeError):
        r._native.render(num_scenes=4, tile_w=64, tile_h=64, channels=3, vp=r._pbr_cam.viewbuf.cpu(),
                         nodes=r._native_nodes(), out=torch.empty((4, 3, 64, 64), dtype=torch.uint8, device="cuda"),
                         bg=(0, 0, 0, 1), ambient=(0, 0, 0), dir_dir=(0, 0, 1), dir_col=(1, 1, 1), strength=1.0)


# ------------------------------------------------------------------ static layer (pbr_base_t)
def _mixed_scene(num_scenes=32, seed=31, tile=(64, 64)):
    """Shared boxes (static layer candidates) interpenetrating per-scene boxes."""
    from pybatchrender_b200 import PBRRenderer
    r = PBRRenderer(dict(num_scenes=num_scenes, tile_resolution=tile, device="cuda"))
    rng = np.random.default_rng(seed)
    shared = r.add_node("models/box", instances_per_scene=2, model_pivot_relative_point=(0.5, 0.5, 0.5),
                        shared_across_scenes=True)
    per = r.add_node("models/box", instances_per_scene=1, model_pivot_relative_point=(0.5, 0.5, 0.5))
    for node in (shared, per):
        B = node.buf_instances
        node.set_positions(torch.tensor(rng.uniform(-2, 2, (B, 3)), dtype=torch.float32), lazy=True)
        node.set_hprs(torch.tensor(rng.uniform(-np.pi, np.pi, (B, 3)), dtype=torch.float32), lazy=True)
        node.set_scales(torch.tensor(rng.uniform(1.0, 3.0, (B, 1)), dtype=torch.float32))
        node.set_colors(torch.tensor(np.concatenate([rng.uniform(0, 1, (B, 3)), np.ones((B, 1))], 1), dtype=torch.float32))
    cam = r.add_camera()
    cam.set_positions(torch.tensor([0.0, -9.0, 1.0]))
    cam.look_at(torch.tensor([0.0, 0.0, 0.0]))
    r.add_light()
    r.setup_environment()
    return r, shared, per


def test_static_layer_is_bit_identical_and_tracks_changes():
    r, shared, per = _mixed_scene()
    assert r._pbr_cam.uniform
    a = r.step()
    assert r._base is not None and r._base_sig is not None          # the static layer was used
    _assert_same(a, oracle_render(r), "static layer")
    r.static_layer = False
    b = r.step()
    assert torch.equal(a, b)
    r.static_layer = True
    # moving a shared instance must re-render the layer
    pos = torch.tensor([[0.5, 0.0, 0.3], [-1.0, 0.5, -0.2]])
    shared.set_positions(pos)
    c = r.step()
    _assert_same(c, oracle_render(r), "static layer after shared update")
    assert not torch.equal(a, c)
    # light and background are part of the layer too
    r._pbr_light.set_ambient((0.5, 0.1, 0.1))
    r.set_background_color(0.2, 0.3, 0.4)
    _assert_same(r.step(), oracle_render(r), "static layer after light/bg update")


def test_static_layer_disabled_by_per_scene_camera():
    r, shared, per = _mixed_scene(num_scenes=8)
    eyes = torch.tensor(np.random.default_rng(2).uniform(-1, 1, (8, 3)), dtype=torch.float32) + torch.tensor([0., -9., 1.])
    r._pbr_cam.set_positions(eyes)
    assert not r._pbr_cam.uniform
    px = r.step()
    assert r._base_sig is None
    _assert_same(px, oracle_render(r), "per-scene camera")
    r._pbr_cam.set_positions(torch.tensor([0.0, -9.0, 1.0]))
    r._pbr_cam.look_at(torch.tensor([0.0, 0.0, 0.0]))
    assert r._pbr_cam.uniform
    _assert_same(r.step(), oracle_render(r), "uniform again")


def test_static_layer_non_square_tiles_and_rgba():
    from pybatchrender_b200.envs.cartpole import CartPoleRenderer
    for tile, ch in [((84, 84), 3), ((50, 30), 4), ((128, 96), 3)]:
        r = CartPoleRenderer(dict(num_scenes=21, tile_resolution=tile, device="cuda", num_channels=ch))
        px = r.step(cartpole_states(21, seed=5).cuda())
        assert r._base_sig is not None
        _assert_same(px, oracle_render(r), f"static layer {tile} C={ch}")


def test_record_overflow_is_exact_and_sticky():
    """Camera inside the geometry: so many triangles are clipped into fans that the small-scene
    kernel runs out of shared-memory record slots.  The surplus goes to its global overflow pool, so
    the frame in flight is still exact; the sticky status bit (also in host-mapped memory) then
    steers later frames on this device to the general kernel, which is exact as well."""
    dev = torch.cuda.current_device()
    overflowed = 0
    for seed in range(40, 60):
        inst = 3 if seed % 2 else 4          # 36 slots (the kernel's maximum) and 48 -> general kernel
        r = many_cubes_renderer(num_scenes=16, instances=inst, tile=(64, 64), device="cuda", seed=seed, spread=0.3,
                                eye=(0.0, 0.0, 0.0), two_sided=True)
        r._native.device_status(dev, clear=True)
        first = r.step()
        status = r._native.device_status(dev, clear=False)
        second = r.step()                      # general kernel if the flag was raised
        ref = oracle_render(r)
        _assert_same(first, ref, f"frame in flight, seed {seed}, status {status}")
        _assert_same(second, ref, f"frame after the overflow check, seed {seed}")
        overflowed += status & 1
    r._native.device_status(dev, clear=True)   # do not leave the device in general-only mode
    assert overflowed >= 3, "too few seeds produced a record overflow (test scene needs adjusting)"


# ------------------------------------------------------------------ BASELINE configs at full size
def test_config4_full_size_65536_scenes_84x84():
    """BASELINE config 4 on one GPU (the multi-GPU run shards the same batch): 65,536 CartPole scenes
    at 84x84, every scene against the oracle."""
    n = 65536
    r = _cartpole(n, (84, 84))
    assert r.cfg.tiles == (256, 256)
    px = r.step(cartpole_states(n, seed=4).cuda())
    assert px.shape == (n, 3, 84, 84)
    _assert_same(px, oracle_render(r, n_threads=host_threads()), "config 4, every scene")
    nonbg = (px != 0).any(1).flatten(1).sum(1)
    assert int(nonbg.min()) > 30 and int(nonbg.max()) < 1200
    # idempotence: rendering the same state again gives the same bytes
    assert torch.equal(px, r.render())


def test_config3_full_size_many_cubes_1024x256_128x128():
    """BASELINE config 3: 1024 scenes x 256 boxes at 128x128 (geometry pre-pass + TMA-staged raster,
    4 bands per tile, ~12 record chunks per scene, near-plane clipping)."""
    r = many_cubes_renderer(num_scenes=1024, instances=256, tile=(128, 128), device="cuda")
    px = r.step()
    assert px.shape == (1024, 3, 128, 128)
    _assert_same(px, oracle_render(r, n_threads=host_threads()), "config 3, every scene")
    assert r._native.device_status(torch.cuda.current_device()) == 0
    assert torch.equal(px, r.render())


def test_sharded_render_equals_single_process_render():
    """Two shard renderers on one GPU (scene_offset / global grid from dist.shard_config) produce the
    same frames as the unsharded renderer."""
    from pybatchrender_b200.dist import shard_config
    from pybatchrender_b200.envs.cartpole import CartPoleConfig, CartPoleRenderer
    n = 300
    st = cartpole_states(n, seed=8).cuda()
    g = CartPoleConfig(num_scenes=n, tile_resolution=(64, 64), device="cuda")
    full = CartPoleRenderer(g).step(st)
    parts = []
    for rank in range(3):
        cfg = shard_config(g, rank, 3)
        parts.append(CartPoleRenderer(cfg).step(st[cfg.scene_offset:cfg.scene_offset + cfg.num_scenes]))
    assert torch.equal(torch.cat(parts), full)


# ------------------------------------------------------------------ smooth shading
@pytest.mark.parametrize("kw", [
    dict(num_scenes=6, boxes=0, spheres=1, segments=8, rings=4),                       # fused general kernel
    dict(num_scenes=6, boxes=4, spheres=3),                                            # staged (TMA) path
    dict(num_scenes=4, boxes=4, spheres=3, channels=4, tile=(96, 80)),
    dict(num_scenes=3, boxes=6, spheres=6, tile=(200, 136), segments=16, rings=12),   # several bands
    dict(num_scenes=4, boxes=2, spheres=4, two_sided=True),
    dict(num_scenes=4, boxes=3, spheres=3, shared_spheres=True),
    dict(num_scenes=4, boxes=2, spheres=6, spread=3.0, eye=(0.0, -2.0, 0.0)),          # heavy near-plane clipping
    dict(num_scenes=4, boxes=0, spheres=3, spread=1.5, eye=(0.2, -0.3, 0.1), two_sided=True),   # camera inside
])
@pytest.mark.parametrize("path", list(LARGE_PATHS))
def test_smooth_normals_bit_exact(kw, path):
    r = mixed_mesh_renderer(device="cuda", **kw)
    r.render_flags = LARGE_PATHS[path]
    _assert_same(r.render(), oracle_render(r), f"smooth {kw} {path}")
    assert r._native.device_status(torch.cuda.current_device()) == 0


def test_config5_mixed_mesh_files_small():
    """Box + models/cone.egg + models/cylinder/scene.gltf + sphere through the file readers, all four
    kinds of normals (flat, smooth, n-gon cap, two-sided glTF material), every byte vs the oracle."""
    r = config5_renderer(num_scenes=3, per_node=16, tile=(256, 256), device="cuda")
    assert [n.mesh.n_tris for n in r._drawable_nodes()][1:3] == [62, 320]
    _assert_same(r.render(), oracle_render(r), "config 5, 3 scenes")
    assert r._native.device_status(torch.cuda.current_device()) == 0


def test_config5_full_size_16384_scenes_256x256():
    """BASELINE config 5 at full size on one GPU: 16,384 scenes x 64 instances at 256x256 (3.2 GB of
    frames, ~12k triangle slots per scene, several staged launches).  Sampled scenes vs the oracle,
    idempotence, no overflow."""
    n = 16384
    r = config5_renderer(num_scenes=n, device="cuda")
    px = r.step()
    assert px.shape == (n, 3, 256, 256)
    # 272 scenes: every 64th (each staged launch holds ~1000 scenes, so every launch is sampled ~16 times)
    # plus the first and last scenes of the batch
    sample = sorted(set(range(0, n, 64)) | set(range(8)) | set(range(n - 8, n)))
    ref = oracle_subset(r, sample, n_threads=host_threads())
    got = px[torch.tensor(sample, device=px.device)]
    _assert_same(got, ref, "config 5, 272 scenes spread over every staged launch")
    assert r._native.device_status(torch.cuda.current_device()) == 0
    again = r.render()
    assert torch.equal(px, again)


@pytest.mark.parametrize("seed,scale", [(0, 0.02), (1, 0.05), (2, 0.15)])
def test_instance_culling_keeps_every_pixel(seed, scale):
    """Staged path with its instance cull (bounding sphere vs clip planes and vs the pixel grid):
    hundreds of spheres around one pixel in size, scattered inside and outside the view, must give
    exactly the oracle's frame (the oracle culls nothing)."""
    from pybatchrender_b200 import PBRRenderer
    r = PBRRenderer(dict(num_scenes=6, tile_resolution=(64, 48), device="cuda"))
    rng = np.random.default_rng(seed)
    node = r.add_node("models/smiley", instances_per_scene=150)
    B = node.buf_instances
    node.set_positions(torch.tensor(rng.uniform(-9, 9, (B, 3)), dtype=torch.float32), lazy=True)
    node.set_hprs(torch.tensor(rng.uniform(-np.pi, np.pi, (B, 3)), dtype=torch.float32), lazy=True)
    node.set_scales(torch.tensor(rng.uniform(0.3, 2.0, (B, 1)) * scale, dtype=torch.float32))
    node.set_colors(torch.tensor(np.concatenate([rng.uniform(0.2, 1, (B, 3)), np.ones((B, 1))], 1), dtype=torch.float32))
    cam = r.add_camera()
    cam.set_positions(torch.tensor([0.0, -12.0, 0.0]))
    r.add_light()
    r.setup_environment()
    got = r.render()
    _assert_same(got, oracle_render(r), f"cull seed={seed} scale={scale}")
    assert int((got != 0).any(1).sum()) > 5


def test_small_scene_kernel_variants_agree_with_the_oracle():
    """The small-scene kernel has two builds -- background written by the warps (<4, false>) or by TMA with
    14 scenes per CTA (<14, true>) -- that the host picks by tile size.  Force each on tile sizes on both
    sides of the automatic choice (the switch is read once per process, hence the subprocesses)."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    script = (
        "import sys, numpy as np, torch\n"
        f"sys.path.insert(0, {os.path.dirname(here)!r}); sys.path.insert(0, {here!r})\n"
        "from util import cartpole_states, oracle_render\n"
        "from pybatchrender_b200.envs.cartpole import CartPoleRenderer\n"
        "for n, tile, ch in ((37, (64, 64), 3), (19, (48, 40), 3), (9, (96, 72), 3), (15, (64, 64), 4), (30, (84, 84), 3), (21, (128, 128), 3)):\n"
        "    r = CartPoleRenderer(dict(num_scenes=n, tile_resolution=tile, device='cuda', num_channels=ch))\n"
        "    got = r.step(cartpole_states(n, seed=n).cuda()).cpu().numpy()\n"
        "    assert np.array_equal(got, oracle_render(r)), (n, tile)\n"
        "    r.static_layer = False\n"
        "    assert np.array_equal(r.render().cpu().numpy(), oracle_render(r)), ('no static layer', n, tile)\n"
        "print('ok')\n")
    for mode in ("0", "1"):
        env = dict(os.environ, PBR_B200_WARP_TMA=mode)
        out = subprocess.run([sys.executable, "-c", script], env=env, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0 and out.stdout.strip().endswith("ok"), f"PBR_B200_WARP_TMA={mode}: {out.stderr[-2000:]}"


def test_frames_on_two_streams_do_not_share_scratch():
    """Two large-scene frames (geometry pre-pass + staged raster, scratch record lists) enqueued back to
    back on different CUDA streams: each stream has its own scratch, both frames equal the oracle."""
    a = many_cubes_renderer(num_scenes=8, instances=40, tile=(96, 96), device="cuda", seed=5)
    b = many_cubes_renderer(num_scenes=8, instances=40, tile=(96, 96), device="cuda", seed=6)
    a.render(); b.render()                       # allocations, static data
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs = []
    for _ in range(3):
        with torch.cuda.stream(s1):
            pa = a.render()
        with torch.cuda.stream(s2):
            pb = b.render()
        outs.append((pa, pb))
    torch.cuda.synchronize()
    ra, rb = oracle_render(a), oracle_render(b)
    for pa, pb in outs:
        _assert_same(pa, ra, "stream 1")
        _assert_same(pb, rb, "stream 2")
