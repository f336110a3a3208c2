"""GPU parity: the CUDA product, called through the C ABI, against the CPU oracle, against the
committed golden frames and against the reference's own kernels (oracle/_ref) on the same inputs.
Bit-exact for paths and colours; shaded colours bit-exact without fog, <= 1/255 with fog (exp/pow)."""
import glob
import os

import numpy as np
import pytest

from conftest import ROOT, get_scene, scene_cameras
from hashdag_b200 import camera
from oracle import hdo

pytestmark = pytest.mark.gpu

W, H = 320, 200


@pytest.fixture(scope="module")
def tr():
    from hashdag_b200 import tracer
    cache = {}

    def get(levels, w=W, h=H):
        if (levels, w, h) not in cache:
            cache[(levels, w, h)] = tracer.DAGTracer(True, w, h, levels)
        return cache[(levels, w, h)]
    yield get
    for t in cache.values():
        t.close()


def _info(s):
    return camera.DAGInfo(s.bounds_min, s.bounds_max)


def _diff(a, b):
    a8 = a.view(np.uint8).astype(np.int16)
    b8 = b.view(np.uint8).astype(np.int16)
    return int(np.abs(a8 - b8).max())


@pytest.mark.parametrize("levels,fp", [(12, 10), (13, 10), (16, 11), (17, 10)])
@pytest.mark.parametrize("kind", ["basic", "hash"])
def test_paths_colors_shadows_match_oracle(tr, levels, fp, kind):
    from hashdag_b200 import tracer
    s = get_scene(levels, fp)
    t = tr(levels)
    hashed = kind == "hash"
    dag = (tracer.HashDAG if hashed else tracer.BasicDAG).from_scene(s)
    odag = hdo.make_dag(s, hdo.DAG_HASH if hashed else hdo.DAG_BASIC)
    with_hash_colors = hashed and s.has_hash_colors
    if hashed and not with_hash_colors:
        col = ocol = None
    else:
        col = (tracer.HashDAGColors if hashed else tracer.BasicDAGCompressedColors).from_scene(s)
        ocol = hdo.make_colors(s, hdo.COLORS_HASH if hashed else hdo.COLORS_COMPRESSED)
    total_hits = 0
    for cam in scene_cameras(s, 3, fp):
        prm = camera.trace_params(cam, _info(s), levels, W, H)
        ms = t.resolve_paths(cam, _info(s), dag)
        assert ms > 0
        p = t.read_paths()
        op, st = hdo.trace_paths(odag, W, H, prm)
        assert np.array_equal(p, op), f"{(p != op).any(-1).sum()} path pixels differ"
        total_hits += st["n_hit"]
        if col is None:
            continue
        t.resolve_colors(dag, col)
        c = t.read_colors()
        oc, _ = hdo.trace_colors(odag, ocol, op)
        assert np.array_equal(c, oc), f"{(c != oc).sum()} colour pixels differ"
        t.resolve_shadows(cam, _info(s), dag, 1.0, 0.0)
        sh = t.read_colors()
        osh, _ = hdo.trace_shadows(odag, prm, op, oc, 1.0, 0.0)
        assert np.array_equal(sh, osh), f"{(sh != osh).sum()} shaded pixels differ"
        t.resolve_colors(dag, col)
        t.resolve_shadows(cam, _info(s), dag, 2.5, 5.0)
        fg = t.read_colors()
        ofg, _ = hdo.trace_shadows(odag, prm, op, oc, 2.5, 5.0)
        assert _diff(fg, ofg) <= 1
        assert (fg != ofg).mean() < 1e-3
    assert total_hits > 0


@pytest.mark.parametrize("dbg", range(1, 8))
def test_debug_color_modes_match_oracle(tr, dbg):
    from hashdag_b200 import tracer
    s = get_scene(13, 10)
    t = tr(13)
    cam = scene_cameras(s, 1, 10)[0]
    prm = camera.trace_params(cam, _info(s), 13, W, H)
    for hashed in (False, True):
        dag = (tracer.HashDAG if hashed else tracer.BasicDAG).from_scene(s)
        col = (tracer.HashDAGColors if hashed else tracer.BasicDAGCompressedColors).from_scene(s)
        odag = hdo.make_dag(s, hdo.DAG_HASH if hashed else hdo.DAG_BASIC)
        ocol = hdo.make_colors(s, hdo.COLORS_HASH if hashed else hdo.COLORS_COMPRESSED)
        t.resolve_paths(cam, _info(s), dag)
        p = t.read_paths()
        for lvl in (0, 4, 11):
            t.resolve_colors(dag, col, dbg, lvl)
            oc, _ = hdo.trace_colors(odag, ocol, p, dbg, lvl)
            assert np.array_equal(t.read_colors(), oc), f"debug mode {dbg} level {lvl} hashed={hashed}"


def test_uncompressed_and_error_colors_match_oracle(tr):
    from hashdag_b200 import tracer
    s = get_scene(12, 10, uncompressed=True)
    t = tr(12)
    cam = scene_cameras(s, 1, 10)[0]
    dag = tracer.BasicDAG.from_scene(s)
    odag = hdo.make_dag(s, hdo.DAG_BASIC)
    t.resolve_paths(cam, _info(s), dag)
    p = t.read_paths()
    unc = tracer.BasicDAGUncompressedColors.from_scene(s)
    comp = tracer.BasicDAGCompressedColors.from_scene(s)
    t.resolve_colors(dag, unc)
    assert np.array_equal(t.read_colors(), hdo.trace_colors(odag, hdo.make_colors(s, hdo.COLORS_UNCOMPRESSED), p)[0])
    t.resolve_colors(dag, tracer.BasicDAGColorErrors(comp, unc))
    assert np.array_equal(t.read_colors(), hdo.trace_colors(odag, hdo.make_colors(s, hdo.COLORS_ERRORS), p)[0])


def test_get_path_and_frame_call(tr):
    from hashdag_b200 import tracer
    s = get_scene(13, 10)
    t = tr(13)
    cam = scene_cameras(s, 1, 10)[0]
    dag, col = tracer.HashDAG.from_scene(s), tracer.HashDAGColors.from_scene(s)
    t.resolve_paths(cam, _info(s), dag)
    p = t.read_paths()
    for (x, y) in ((0, 0), (W // 2, H // 2), (W - 1, H - 1), (17, 133)):
        assert t.get_path(x, y) == tuple(int(v) for v in p[y, x, :3])
    t.resolve_colors(dag, col)
    t.resolve_shadows(cam, _info(s), dag, 1.0, 0.0)
    want = t.read_colors()
    host = np.zeros((H, W), dtype=np.uint32)
    ms = t.resolve_frame(cam, _info(s), dag, col, 1.0, 0.0, True, host)
    assert all(m > 0 for m in ms)
    assert np.array_equal(host, want) and np.array_equal(t.read_paths(), p)


def test_abi_rejects_bad_arguments(tr):
    from hashdag_b200 import tracer
    s = get_scene(12, 10)
    t = tr(12)
    dag = tracer.BasicDAG.from_scene(s)
    cam = scene_cameras(s, 1, 10)[0]

    class Short:
        kind = tracer.DAG_BASIC

        def pod(self):
            return dag.pod()[:8]
    with pytest.raises(tracer.TracerError):
        t.resolve_paths(cam, _info(s), Short())
    with pytest.raises(tracer.TracerError):
        t.get_path(W, 0)
    with pytest.raises(tracer.TracerError):   # HashDAG colours with a BasicDAG: not an instantiation the tracer has
        t.resolve_colors(dag, tracer.HashDAGColors(dag.data, dag.data, tracer.CompressedColorLeaf(None, None, None)))
    with pytest.raises(tracer.TracerError):
        tracer.DAGTracer(True, 64, 64, 40)


def test_partitioned_render_assembles_to_the_whole_frame(tr):
    """Screen-tile partition (rank r of n renders tiles t % n == r): gathered == unpartitioned frame."""
    import torch
    from hashdag_b200 import tracer
    s = get_scene(13, 10)
    cam = scene_cameras(s, 1, 10)[0]
    dag, col = tracer.HashDAG.from_scene(s), tracer.HashDAGColors.from_scene(s)
    whole = tr(13)
    whole.resolve_frame(cam, _info(s), dag, col, 1.0, 0.0, True)
    want_c, want_p = whole.read_colors(), whole.read_paths()
    world, tl = 3, 5
    parts = []
    acc_c = np.zeros_like(want_c)
    acc_p = np.zeros_like(want_p)
    for r in range(world):
        t = tracer.DAGTracer(True, W, H, 13)
        t.set_partition(r, world, tl)
        t.resolve_frame(cam, _info(s), dag, col, 1.0, 0.0, True)
        acc_c |= t.read_colors()
        acc_p |= t.read_paths()
        _, cptr, n_owned, max_tiles = t.partition_buffers()
        n = max_tiles << (2 * tl)
        buf = torch.empty(n, dtype=torch.int32, device="cuda")
        import ctypes
        torch.cuda.synchronize()
        assert ctypes.CDLL("libcudart.so").cudaMemcpy(ctypes.c_void_p(buf.data_ptr()), ctypes.c_void_p(cptr), ctypes.c_size_t(n * 4), 3) == 0
        parts.append(buf)
        t.close()
    assert np.array_equal(acc_c, want_c) and np.array_equal(acc_p, want_p)
    gathered = torch.cat(parts)
    t0 = tracer.DAGTracer(True, W, H, 13)
    t0.set_partition(0, world, tl)
    frame = torch.zeros(H * W, dtype=torch.int32, device="cuda")
    t0.assemble_colors(gathered, frame)
    t0.sync()
    assert np.array_equal(frame.cpu().numpy().view(np.uint32).reshape(H, W), want_c)
    # host mirror of the same layout
    from hashdag_b200 import partition
    packed = [partition.pack_compact(want_c, r, world, tl) for r in range(world)]
    assert np.array_equal(partition.assemble(packed, W, H, tl), want_c)
    assert np.array_equal(packed[1].reshape(-1), parts[1].cpu().numpy().view(np.uint32))
    t0.close()


@pytest.mark.parametrize("levels", [13, 17])
def test_resolved_hash_dag_gives_the_same_frames(tr, levels):
    """HDT_DAG_HASH_RESOLVED (child pointers pre-translated, csrc/hdt_resolve.cuh) against the plain HashDAG: paths,
    every colour view, shaded and fogged frames, beams on and off, and the region queries."""
    import torch
    from hashdag_b200 import tracer
    s = get_scene(levels, 10)
    t = tr(levels)
    dag, col = tracer.HashDAG.from_scene(s), tracer.HashDAGColors.from_scene(s)
    res = t.resolve_hash_dag(dag)
    assert res.kind == tracer.DAG_HASH_RESOLVED and len(res.pod()) == 48 and res.prefix_pool is not None
    info = _info(s)
    # the resolved pool differs from the pool exactly in the child-pointer words
    t.sync()
    a, b = dag.pool.cpu().numpy().view(np.uint32), res.resolved_pool.cpu().numpy().view(np.uint32)
    assert a.shape == b.shape and 0 < int((a != b).sum()) < a.size
    if levels == 13:   # word for word the host mirror of the kernel (itself checked against a walk of the DAG in test_edits_cpu.py)
        from hashdag_b200 import edits
        top = int(s.hash_pool_top)
        assert np.array_equal(b[: top * 512], edits.resolve_pool_host(edits.HashLayout(levels), s.hash_pool, s.hash_page_table, top))
    for beams in (1, 0):
        t.set_option(tracer.OPT_BEAMS, beams)
        for cam in scene_cameras(s, 2, 10) + _special_cameras(s, 10)[:2]:
            frames = []
            for d in (dag, res):
                t.resolve_paths(cam, info, d)
                p = t.read_paths()
                views = []
                for dbg, lvl in ((tracer.DEBUG_NONE, 0), (tracer.DEBUG_INDEX, 3), (tracer.DEBUG_INDEX, levels - 2), (tracer.DEBUG_POSITION, 0),
                                 (tracer.DEBUG_COLOR_TREE, 0), (tracer.DEBUG_WEIGHT, 0)):
                    t.resolve_colors(d, col, dbg, lvl)
                    views.append(t.read_colors())
                t.resolve_colors(d, col)
                t.resolve_shadows(cam, info, d, 1.0, 0.0)
                sh = t.read_colors()
                t.resolve_colors(d, col)
                t.resolve_shadows(cam, info, d, 1.0, 4.0)
                frames.append((p, views, sh, t.read_colors()))
            (p0, v0, s0, f0), (p1, v1, s1, f1) = frames
            assert np.array_equal(p0, p1) and np.array_equal(s0, s1) and np.array_equal(f0, f1)
            for x, y in zip(v0, v1):
                assert np.array_equal(x, y)
    t.set_option(tracer.OPT_BEAMS, 1)
    c = 1 << (levels - 1)
    h0 = int(s.heights[(c, c)])
    st, sz = (c - 20, h0 - 16, c - 20), (50, 40, 45)
    assert torch.equal(t.get_values(dag, st, sz)[0], t.get_values(res, st, sz)[0])
    for lvl in range(levels - 1):
        assert t.is_empty(dag, lvl, st, sz) == t.is_empty(res, lvl, st, sz)


class _DevMem:
    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<i4", "data": (ptr, False), "version": 2}


@pytest.mark.parametrize("fused", [0, 1])
def test_peer_exchange_assembles_the_whole_frame(tr, fused):
    """fused = 1: trace_shadows_kernel stores the final colours into rank 0's frame itself (HDT_OPT_EXCHANGE_FUSED).
    hdt_exchange_*: three ranks (contexts of this process, attached by pointer) store their tiles straight into rank
    0's frame; arrivals / credit counters order the frames.  Three frames in a row == the unpartitioned frames."""
    import torch
    from hashdag_b200 import tracer
    s = get_scene(13, 10)
    cams = scene_cameras(s, 3, 10)[:3]
    dag, col = tracer.HashDAG.from_scene(s), tracer.HashDAGColors.from_scene(s)
    whole = tr(13)
    world, tl = 3, 5
    ranks = []
    for r in range(world):
        t = tracer.DAGTracer(True, W, H, 13)
        t.set_partition(r, world, tl)
        ranks.append(t)
    _, frame_ptr = ranks[0].exchange_create()
    for t in ranks[1:]:
        t.exchange_attach(ranks[0].exchange_block())
    for t in ranks:
        t.set_option(tracer.OPT_EXCHANGE_FUSED, fused)
    frame = torch.as_tensor(_DevMem(frame_ptr, W * H), device="cuda")
    info = _info(s)
    for cam in cams:
        whole.resolve_frame(cam, info, dag, col, 1.0, 0.0, True)
        want = whole.read_colors()
        prm = camera.trace_params(cam, info, 13, W, H)
        for t in ranks[1:] + ranks[:1]:          # the root last: its wait needs the peers' work queued (one host thread here)
            t.enqueue_frame(prm, dag.pod(), dag.kind, col.pod(), col.kind, 1.0, 0.0, True, None)
            t.exchange_frame()
        ranks[0].sync()
        got = frame.cpu().numpy().view(np.uint32).reshape(H, W)
        assert np.array_equal(got, want)
        frame.zero_()
        torch.cuda.synchronize()
        ranks[0].exchange_release()
    for t in ranks:
        t.sync()
        t.close()


def test_peer_exchange_gives_up_on_a_missing_rank():
    from hashdag_b200 import tracer
    s = get_scene(13, 10)
    t = tracer.DAGTracer(True, W, H, 13)
    t.set_partition(0, 2, 5)
    t.exchange_create()
    t.set_option(tracer.OPT_EXCHANGE_TIMEOUT_MS, 300)
    t.exchange_frame()                            # rank 1 never arrives
    with pytest.raises(tracer.TracerError):
        t.sync()
    t.sync()                                      # the error is reported once
    t.close()


GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "ref_d[0-9]*.npz")))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_cuda_against_reference_golden_frames(path):
    from golden_util import check_golden
    check_golden(path, impl="cuda")


def _special_cameras(s, fp):
    """Poses that stress the beam pre-pass: outside the volume (root misses and partial root hits),
    axis-parallel views (non-tame rays: no beam), direction sign changes inside a tile, grazing views."""
    c = float(1 << (s.levels - 1))
    size = float(1 << s.levels)
    h0 = float(s.heights.get((int(c), int(c)), c))
    r = float(1 << fp)
    return [
        camera.look_at((-0.3 * size, 1.2 * size, -0.2 * size), (c, h0, c)),                  # far outside, volume in view
        camera.look_at((c - 2.0 * r, h0 + 0.1 * r, c), (c + r, h0 + 0.1 * r, c + 1e-3)),     # grazing, nearly axis-parallel
        camera.CameraView((c, h0 + 0.6 * r, c), ((1.0, 0.0, 0.0), (0.0, 0.0, 1.0), (0.0, -1.0, 0.0))),  # straight down
        camera.CameraView((c + 0.25, h0 + 30.5, c + 0.75), ((1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0))),  # along +z, signs flip mid-frame
        camera.look_at((c + 7.3, h0 + 3.1, c - 5.2), (c + 40.0, h0 - 6.0, c + 33.0)),         # a few voxels above the ground
    ]


@pytest.mark.parametrize("levels,fp,w,h", [(13, 10, 320, 200), (17, 11, 333, 203), (16, 11, 640, 360)])
@pytest.mark.parametrize("kind", ["basic", "hash"])
def test_beam_prepass_never_changes_a_pixel(levels, fp, w, h, kind):
    """hdt_beam.cuh: frames with the per-tile beam pre-pass == frames without == oracle, and the
    pre-pass is really exercised (tiles resume below the root)."""
    from hashdag_b200 import tracer
    s = get_scene(levels, fp)
    hashed = kind == "hash"
    if hashed and not s.has_hash_colors:
        pytest.skip("no hash colours at this depth")
    t = tracer.DAGTracer(True, w, h, levels)
    dag = (tracer.HashDAG if hashed else tracer.BasicDAG).from_scene(s)
    col = (tracer.HashDAGColors if hashed else tracer.BasicDAGCompressedColors).from_scene(s)
    odag = hdo.make_dag(s, hdo.DAG_HASH if hashed else hdo.DAG_BASIC)
    ocol = hdo.make_colors(s, hdo.COLORS_HASH if hashed else hdo.COLORS_COMPRESSED)
    resumed = 0
    for cam in scene_cameras(s, 2, fp) + _special_cameras(s, fp):
        frames = {}
        for beams in (1, 0):
            t.set_option(tracer.OPT_BEAMS, beams)
            t.resolve_paths(cam, _info(s), dag)
            if beams:
                st = t.beam_stats()
                assert st["root"] + st["resume"] + st["hit"] + st["miss"] >= ((w + 7) // 8) * ((h + 3) // 4)
                resumed += st["resume"] + st["hit"] + st["miss"]
            p = t.read_paths()
            t.resolve_colors(dag, col)
            t.resolve_shadows(cam, _info(s), dag, 1.0, 0.0)
            frames[beams] = (p, t.read_colors())
        assert np.array_equal(frames[1][0], frames[0][0]), f"{(frames[1][0] != frames[0][0]).any(-1).sum()} path pixels change with beams"
        assert np.array_equal(frames[1][1], frames[0][1]), f"{(frames[1][1] != frames[0][1]).sum()} shaded pixels change with beams"
        prm = camera.trace_params(cam, _info(s), levels, w, h)
        op, _ = hdo.trace_paths(odag, w, h, prm)
        assert np.array_equal(frames[1][0], op)
        oc, _ = hdo.trace_colors(odag, ocol, op)
        osh, _ = hdo.trace_shadows(odag, prm, op, oc, 1.0, 0.0)
        assert np.array_equal(frames[1][1], osh)
    assert resumed > 0, "the beam pre-pass never took over a tile"
    t.close()


def test_replay_file_drives_the_tracer_like_the_engine(tr, tmp_path):
    """hashdag_b200/replay.py: a camera track written in the reference's replay CSV format, read back and run
    through resolve_paths / resolve_colors / resolve_shadows per frame; stats rows as the reference reports them."""
    from hashdag_b200 import replay, tracer
    s = get_scene(13, 10)
    t = tr(13)
    dag, col = tracer.HashDAG.from_scene(s), tracer.HashDAGColors.from_scene(s)
    cams = scene_cameras(s, 3, 10)[:4]
    path = str(tmp_path / "track.csv")
    replay.dump([replay.Frame(c) for c in cams], path)
    frames = replay.load(path)
    assert len(frames) == 4
    stats = replay.run(t, frames, _info(s), dag, col, 1.0, 0.0)
    last = t.read_colors()
    rows = replay.StatsRecorder.read_csv(stats.to_csv(str(tmp_path / "track.stats.csv")))
    assert [r[1] for r in rows] == ["paths", "colors", "shadows"] * 4 and all(r[2] > 0 for r in rows) and rows[-1][0] == 3
    # std::to_string keeps 6 decimals: the replayed camera is the written one rounded, and the frame is its frame
    cam = frames[-1].camera
    t.resolve_paths(cam, _info(s), dag); t.resolve_colors(dag, col); t.resolve_shadows(cam, _info(s), dag, 1.0, 0.0)
    assert np.array_equal(t.read_colors(), last)
    odag, ocol = hdo.make_dag(s, hdo.DAG_HASH), hdo.make_colors(s, hdo.COLORS_HASH)
    prm = camera.trace_params(cam, _info(s), 13, W, H)
    op, _ = hdo.trace_paths(odag, W, H, prm)
    oc, _ = hdo.trace_colors(odag, ocol, op)
    osh, _ = hdo.trace_shadows(odag, prm, op, oc, 1.0, 0.0)
    assert np.array_equal(last, osh)


def _reachable_words(scene):
    """Physical word indices of every word a walk of the HashDAG from its root can touch (headers, child pointers, leaves)."""
    pool, table = scene.hash_pool, scene.hash_page_table.astype(np.int64)
    phys = lambda v: table[v >> 9] * 512 + (v & 511)
    used = []
    frontier = np.array([scene.hash_first_node_index], dtype=np.int64)
    for level in range(scene.levels - 2):
        h = phys(frontier)
        hdr = pool[h].astype(np.int64)
        n = np.array([bin(int(x) & 0xFF).count("1") for x in np.unique(hdr & 0xFF)])   # popcount table of the masks present
        lut = np.zeros(256, np.int64)
        lut[np.unique(hdr & 0xFF)] = n
        cnt = lut[hdr & 0xFF]
        idx = np.repeat(h, cnt) + (np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt)) + 1
        used += [h, idx]
        frontier = np.unique(pool[idx].astype(np.int64))
    h = phys(frontier)
    used += [h, h + 1]
    return np.unique(np.concatenate(used))


@pytest.mark.parametrize("levels,fp", [(13, 10), (16, 11), (17, 10)])
def test_recorded_colors_equal_the_full_walk(tr, levels, fp):
    """HDT_OPT_COLORS_RECORDED: trace_paths leaves per hit pixel where its path leaves each ancestor below the colour tree,
    trace_colors reads the colour index off the prefix pool.  Same pixels as the walk of tracer.cu:300-430 (option off) and
    as the oracle: every decoding view, the tool overlay, beams on and off, hostile cameras."""
    from hashdag_b200 import tracer
    s = get_scene(levels, fp)
    t = tr(levels)
    dag, col = tracer.HashDAG.from_scene(s), tracer.HashDAGColors.from_scene(s)
    res = t.resolve_hash_dag(dag)
    odag, ocol = hdo.make_dag(s, hdo.DAG_HASH), hdo.make_colors(s, hdo.COLORS_HASH)
    info = _info(s)
    c = 1 << (levels - 1)
    tool = tracer.ToolInfo(0, (c, int(s.heights[(c, c)]), c), 60.0, (0, 0, 0), (0, 0, 0))
    views = [(tracer.DEBUG_NONE, False), (tracer.DEBUG_COLOR_BITS, False), (tracer.DEBUG_MIN_COLOR, False), (tracer.DEBUG_MAX_COLOR, False),
             (tracer.DEBUG_WEIGHT, False), (tracer.DEBUG_NONE, True)]
    for beams in (1, 0):
        t.set_option(tracer.OPT_BEAMS, beams)
        for cam in scene_cameras(s, 2, fp) + _special_cameras(s, fp)[:3]:
            frames = {}
            for recorded in (1, 0):
                t.set_option(tracer.OPT_COLORS_RECORDED, recorded)
                before = t.recorded_color_passes()
                t.resolve_paths(cam, info, res)
                frames[recorded] = []
                for dbg, overlay in views:
                    t.resolve_colors(res, col, dbg, 0, tool if overlay else None, overlay)
                    frames[recorded].append(t.read_colors())
                t.resolve_colors(res, col, tracer.DEBUG_INDEX, 3)          # a view the records cannot serve: full walk
                assert t.recorded_color_passes() - before == (len(views) if recorded else 0)
            for a, b in zip(frames[1], frames[0]):
                assert np.array_equal(a, b), f"{(a != b).sum()} pixels differ between the recorded route and the walk"
            oc, _ = hdo.trace_colors(odag, ocol, t.read_paths())
            assert np.array_equal(frames[1][0], oc)
    t.set_option(tracer.OPT_BEAMS, 1)
    t.set_option(tracer.OPT_COLORS_RECORDED, 1)
    # a colours pass for ANOTHER DAG than the paths frame was traced in must not use the records
    t.resolve_paths(scene_cameras(s, 1, fp)[0], info, res)
    before = t.recorded_color_passes()
    t.resolve_colors(dag, col)
    other = t.resolve_hash_dag(dag)
    t.resolve_colors(other, col)
    assert t.recorded_color_passes() == before
    t.resolve_colors(res, col)
    assert t.recorded_color_passes() == before + 1


def test_resolve_ignores_what_lies_behind_the_nodes(tr):
    """The reference never clears its pool (cudaMalloc'd, hash_table.cpp:60-76): words no node occupies may hold anything.
    With every such word set to 0xFFFFFFFF the resolved and prefix pools must still serve the same frames."""
    import torch
    from hashdag_b200 import tracer
    levels = 13
    s = get_scene(levels, 10)
    t = tr(levels)
    live = _reachable_words(s)
    dirty = np.full(s.hash_pool.size + 3 * 512, 0xFFFFFFFF, dtype=np.uint32)       # garbage pages behind pool_top too
    dirty[live] = s.hash_pool[live]
    assert (dirty[: s.hash_pool.size] != s.hash_pool).sum() > 1000
    clean, col = tracer.HashDAG.from_scene(s), tracer.HashDAGColors.from_scene(s)
    bad = tracer.HashDAG(tracer._to_device(dirty, "cuda:0"), clean.page_table, clean.pool_top, clean.first_node_index, levels)
    res_clean, res_bad = t.resolve_hash_dag(clean), t.resolve_hash_dag(bad)
    t.sync()
    info = _info(s)
    for cam in scene_cameras(s, 2, 10):
        out = []
        for d in (res_clean, res_bad, bad):
            t.resolve_paths(cam, info, d)
            p = t.read_paths()
            t.resolve_colors(d, col)
            t.resolve_shadows(cam, info, d, 1.0, 0.0)
            out.append((p, t.read_colors()))
        for p, c in out[1:]:
            assert np.array_equal(p, out[0][0]) and np.array_equal(c, out[0][1])
    # live words of the two resolved / prefix pools agree
    a, b = res_clean.resolved_pool.cpu().numpy().view(np.uint32), res_bad.resolved_pool.cpu().numpy().view(np.uint32)
    assert np.array_equal(a[live], b[live])
    a, b = res_clean.prefix_pool.cpu().numpy().view(np.uint32), res_bad.prefix_pool.cpu().numpy().view(np.uint32)
    assert np.array_equal(a[live], b[live])
    torch.cuda.synchronize()
