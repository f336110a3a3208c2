"""CPU: the reference's replay / stats file formats (hashdag_b200/replay.py vs replay.h:76-635, stats.h:35-68)."""
import glob
import os

import pytest

from hashdag_b200 import replay
from hashdag_b200.camera import CameraView

SAMPLE = """SetLocation,-13076.171792,-1671.669064,5849.330320
SetRotation,-0.573465,0.000000,-0.819230,-0.034067,0.999135,0.023847,0.818522,0.041585,-0.572969
EndFrame
SetLocation,-13070.000000,-1671.500000,5850.250000
SetRotation,-0.573465,0.000000,-0.819230,-0.034067,0.999135,0.023847,0.818522,0.041585,-0.572969
SetToolParameters,12385,33035,24094,0,0,0,0,0,0,200.000000,0
EditSphere,12385.000000,33035.000000,24094.000000,200.000000,true
EndFrame
SetLocation,-13070.000000,-1671.500000,5850.250000
SetRotation,-0.573465,0.000000,-0.819230,-0.034067,0.999135,0.023847,0.818522,0.041585,-0.572969
EditCopy,1.000000,2.000000,3.000000,4.000000,5.000000,6.000000,7.000000,8.000000,9.000000,32.000000,1.000000,0.000000,0.000000,0.000000,1.000000,0.000000,0.000000,0.000000,1.000000,false,1.000000
EditCube,10.000000,20.000000,30.000000,5.000000,false
EditFill,1.500000,2.500000,3.500000,8.000000
EditPaint,1.500000,2.500000,3.500000,8.000000
Undo
Redo
EndFrame
"""


def test_load_gives_frames_cameras_and_edits():
    frames = replay.load(SAMPLE.splitlines())
    assert len(frames) == 3
    assert frames[0].camera.position == (-13076.171792, -1671.669064, 5849.33032) and not frames[0].other
    assert frames[0].camera.rotation[1] == (-0.034067, 0.999135, 0.023847)
    assert [a.kind for a in frames[1].other] == ["SetToolParameters", "EditSphere"]
    assert frames[1].other[0].values == (12385, 33035, 24094, 0, 0, 0, 0, 0, 0, 200.0, 0)
    assert frames[1].edits[0].values == (12385.0, 33035.0, 24094.0, 200.0, True)
    assert [a.kind for a in frames[2].edits] == ["EditCopy", "EditCube", "EditFill", "EditPaint", "Undo", "Redo"]
    assert frames[2].edits[0].values[19] is False and len(frames[2].edits[0].values) == 21
    # forward = third row of the rotation (camera_view.h:21-27)
    assert frames[1].camera.forward() == (0.818522, 0.041585, -0.572969)


def test_dump_is_the_inverse_of_load():
    assert replay.dump(replay.load(SAMPLE.splitlines())) == SAMPLE


def test_malformed_rows_are_refused():
    for bad in ("EditSphere,1,2,3,4", "SetRotation,1,2,3", "Teleport,1,2,3", "EditCube,1,2,3,4,maybe"):
        with pytest.raises(ValueError):
            replay.load([bad, "EndFrame"])


@pytest.mark.parametrize("path", sorted(glob.glob("/root/reference/replays/*.csv")), ids=os.path.basename)
def test_every_shipped_replay_round_trips(path):
    """Where the reference is mounted (this container): all of its own replay files parse, and writing them
    back reproduces them byte for byte (modulo the letter case of booleans and blank cells)."""
    text = open(path).read()
    frames = replay.load(path)
    assert len(frames) == text.count("EndFrame")
    norm = [",".join(c for c in l.strip().split(",") if c != "").replace("TRUE", "true").replace("FALSE", "false") for l in text.splitlines() if l.strip()]
    out = replay.dump(frames).splitlines()
    assert len(out) == len(norm)
    for a, b in zip(norm, out):
        if a != b:      # hand-edited files carry more digits than std::to_string writes: same numbers to 6 decimals
            ca, cb = a.split(","), b.split(",")
            assert ca[0] == cb[0] and len(ca) == len(cb)
            assert all(x == y or abs(float(x) - float(y)) <= 5.0e-7 for x, y in zip(ca[1:], cb[1:])), (a, b)
    again = replay.load(out)
    assert [(f.camera, f.actions) for f in again] == [(f.camera, [replay._parse(x.kind, x.to_row().split(",")[1:]) for x in f.actions]) for f in frames]


def test_fit_to_scene_keeps_the_motion_inside_the_bounds():
    frames = replay.load(SAMPLE.splitlines())
    fitted = replay.fit_to_scene(frames, (0.0, 0.0, 0.0), (4096.0, 4096.0, 4096.0))
    for a, b in zip(frames, fitted):
        assert a.camera.rotation == b.camera.rotation and a.other == b.actions
        assert all(0.0 <= v <= 4096.0 for v in b.camera.position)
    d0 = [q - p for p, q in zip(frames[0].camera.position, frames[1].camera.position)]
    d1 = [q - p for p, q in zip(fitted[0].camera.position, fitted[1].camera.position)]
    s = d1[0] / d0[0]
    assert s > 0 and all(abs(y - s * x) < 1e-6 * abs(s) + 1e-9 for x, y in zip(d0, d1))


def test_stats_csv_matches_the_reference_layout():
    st = replay.StatsRecorder()
    st.report("paths", 0.512); st.report("colors", 0.25); st.report("shadows", 0.125); st.next_frame()
    st.report("paths", 1.5); st.next_frame()
    text = st.to_csv()
    assert text == "0,paths,0.512\n0,colors,0.25\n0,shadows,0.125\n1,paths,1.5\n"
    assert replay.StatsRecorder.read_csv(text) == [(0, "paths", 0.512), (0, "colors", 0.25), (0, "shadows", 0.125), (1, "paths", 1.5)]


def test_run_drives_the_three_passes_per_frame():
    calls = []

    class FakeTracer:
        def resolve_paths(self, cam, info, dag): calls.append(("p", cam.position, dag)); return 1.0
        def resolve_colors(self, dag, col): calls.append(("c", dag, col)); return 2.0
        def resolve_shadows(self, cam, info, dag, bias, fog): calls.append(("s", bias, fog)); return 3.0
    frames = replay.load(SAMPLE.splitlines())
    seen = []
    st = replay.run(FakeTracer(), frames, None, "dag0", "col0", 1.0, 0.0, on_edit=lambda a: seen.append(a.kind) or (("dag1", "col1") if a.kind == "EditSphere" else None))
    assert [c[0] for c in calls] == ["p", "c", "s"] * 3
    assert calls[0][2] == "dag0" and calls[3][2] == "dag1"          # the edit of frame 1 is applied before frame 1 is traced
    assert seen == ["EditSphere", "EditCopy", "EditCube", "EditFill", "EditPaint", "Undo", "Redo"]
    assert st.to_csv().splitlines()[:4] == ["0,paths,1", "0,colors,2", "0,shadows,3", "1,paths,1"]
