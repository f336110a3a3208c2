"""The reference's replay and stats files, so its recorded workloads drive this tracer unchanged.

Replay CSV (/root/reference/src/replay.h:76-635, replays/*.csv): one action per line,
`<Action>,<fields...>`; a frame is everything up to and including `EndFrame`.  Camera actions
(`SetLocation x,y,z`, `SetRotation` 9 row-major floats = CameraView::rotation) feed resolve_*;
tool / edit actions (`SetToolParameters`, `EditSphere`, `EditCube`, `EditCopy`, `EditFill`, `EditPaint`,
`Undo`, `Redo`) are parsed and kept for the host editor (edits are outside the traced path).
Stats CSV (stats.h:35-68): `frame,name,value` per line, as python/tools.py of the reference reads it.
"""
from __future__ import annotations

from dataclasses import dataclass, field

from .camera import CameraView

# action -> number(s) of fields it carries (replay.h load() methods)
FIELDS = {"EndFrame": (0,), "Undo": (0,), "Redo": (0,), "SetLocation": (3,), "SetRotation": (9,), "SetToolParameters": (11,),
          "EditSphere": (5,), "EditCube": (5,), "EditCopy": (10, 21), "EditFill": (4,), "EditPaint": (4,)}


def _f(x: float) -> str:
    return f"{float(x):.6f}"          # std::to_string(float/double)


def _b(s: str) -> bool:
    if s not in ("true", "false", "TRUE", "FALSE"):
        raise ValueError(f"replay: expected true/false, got {s!r}")
    return s in ("true", "TRUE")


@dataclass
class Action:
    kind: str
    values: tuple = ()

    def to_row(self) -> str:
        out = [self.kind]
        for v in self.values:
            out.append(("true" if v else "false") if isinstance(v, bool) else str(v) if isinstance(v, int) else _f(v))
        return ",".join(out)


@dataclass
class Frame:
    camera: CameraView | None = None                 # the view when the frame ends (camera actions persist across frames)
    actions: list = field(default_factory=list)      # every action of the frame but EndFrame, in file order

    @property
    def edits(self):
        return [a for a in self.actions if a.kind.startswith("Edit") or a.kind in ("Undo", "Redo")]

    @property
    def other(self):
        return [a for a in self.actions if a.kind not in ("SetLocation", "SetRotation")]


def _parse(kind: str, cells: list) -> Action:
    if kind not in FIELDS:
        raise ValueError(f"replay: unknown action {kind!r}")
    if len(cells) not in FIELDS[kind]:
        raise ValueError(f"replay: {kind} carries {FIELDS[kind]} fields, got {len(cells)}")
    if kind == "SetToolParameters":
        vals = tuple(int(c) for c in cells[:9]) + (float(cells[9]), int(cells[10]))
    elif kind in ("EditSphere", "EditCube"):
        vals = tuple(float(c) for c in cells[:4]) + (_b(cells[4]),)
    elif kind == "EditCopy" and len(cells) == 21:
        vals = tuple(float(c) for c in cells[:19]) + (_b(cells[19]), float(cells[20]))
    else:
        vals = tuple(float(c) for c in cells)
    return Action(kind, vals)


def load(path_or_lines) -> list:
    """-> [Frame].  Camera state persists across frames like Engine::view does (replay.cpp apply())."""
    lines = open(path_or_lines).read().splitlines() if isinstance(path_or_lines, str) else list(path_or_lines)
    frames, cur = [], Frame()
    pos, rot = None, None
    for n, line in enumerate(lines, 1):
        cells = [c for c in line.strip().split(",") if c != ""]        # "remove empty cells", replay.h:560
        if not cells:
            continue
        a = _parse(cells[0], cells[1:])
        if a.kind == "SetLocation":
            pos = a.values
            cur.actions.append(a)
        elif a.kind == "SetRotation":
            rot = (a.values[0:3], a.values[3:6], a.values[6:9])
            cur.actions.append(a)
        elif a.kind == "EndFrame":
            if pos is not None and rot is not None:
                cur.camera = CameraView(tuple(pos), tuple(tuple(r) for r in rot))
            frames.append(cur)
            cur = Frame()
        else:
            cur.actions.append(a)
    return frames


def dump(frames, path=None) -> str:
    """The inverse of load(): the frame's actions in their order, then EndFrame.  A frame built in code (a camera
    but no camera actions) gets its SetLocation / SetRotation first, like ReplayManager records them."""
    rows = []
    for fr in frames:
        if fr.camera is not None and not any(a.kind in ("SetLocation", "SetRotation") for a in fr.actions):
            rows.append(Action("SetLocation", tuple(fr.camera.position)).to_row())
            rows.append(Action("SetRotation", tuple(v for r in fr.camera.rotation for v in r)).to_row())
        rows += [a.to_row() for a in fr.actions]
        rows.append("EndFrame")
    text = "\n".join(rows) + "\n"
    if path:
        open(path, "w").write(text)
    return text


def fit_to_scene(frames, bounds_min, bounds_max, margin=0.08, ground=None, altitude=None):
    """The shipped replays were recorded in Epic Citadel's world frame; its scene file (and DAGInfo) is not
    available offline.  Map the camera track affinely (one uniform scale, x/z centred, rotations untouched)
    into the given bounds so the same motion flies over another scene.  With `ground(x, z) -> height` and
    `altitude`, the track's height above its own minimum is kept above the terrain instead."""
    cams = [f.camera for f in frames if f.camera is not None]
    if not cams:
        return frames
    lo = [min(c.position[k] for c in cams) for k in range(3)]
    hi = [max(c.position[k] for c in cams) for k in range(3)]
    size = [bounds_max[k] - bounds_min[k] for k in range(3)]
    span = max(max(hi[k] - lo[k] for k in (0, 2)), 1e-9)
    s = min(size[0], size[2]) * (1 - 2 * margin) / span
    out = []
    for f in frames:
        if f.camera is None:
            out.append(f)
            continue
        p = f.camera.position
        x = bounds_min[0] + size[0] / 2 + (p[0] - (lo[0] + hi[0]) / 2) * s
        z = bounds_min[2] + size[2] / 2 + (p[2] - (lo[2] + hi[2]) / 2) * s
        y = bounds_min[1] + size[1] / 2 + (p[1] - (lo[1] + hi[1]) / 2) * s
        if ground is not None:
            y = ground(x, z) + (altitude if altitude is not None else 0.0) + (p[1] - lo[1]) * s
        out.append(Frame(CameraView((x, y, z), f.camera.rotation), f.other))
    return out


class StatsRecorder:
    """stats.h:15-101: (frame, name, value) rows; `report` in the current frame, `next_frame` after EndFrame."""

    def __init__(self):
        self.frame, self.elements = 0, []

    def report(self, name: str, value: float):
        self.elements.append((self.frame, name, float(value)))

    def next_frame(self):
        self.frame += 1

    def clear(self):
        self.frame, self.elements = 0, []

    def to_csv(self, path=None) -> str:
        text = "".join(f"{f},{n},{v:g}\n" for f, n, v in self.elements)      # ostream << double: %g
        if path:
            open(path, "w").write(text)
        return text

    @staticmethod
    def read_csv(path_or_text):
        text = open(path_or_text).read() if "\n" not in path_or_text else path_or_text
        rows = []
        for line in text.splitlines():
            if line.strip():
                f, n, v = line.split(",")
                rows.append((int(f), n, float(v)))
        return rows


def run(tracer_obj, frames, dag_info, dag, colors, shadow_bias=1.0, fog_density=0.0, stats: StatsRecorder | None = None, on_edit=None):
    """Engine::tick over a replay (engine.cpp:575-648, 748): per frame the three resolve_* calls with the frame's
    camera, their times reported as "paths" / "colors" / "shadows" like the reference does; edit actions go to
    `on_edit(action) -> (dag, colors) | None` (the host editor) before the frame is traced."""
    stats = stats or StatsRecorder()
    for fr in frames:
        if on_edit:
            for a in fr.edits:
                res = on_edit(a)
                if res:
                    dag, colors = res
        if fr.camera is not None:
            stats.report("paths", tracer_obj.resolve_paths(fr.camera, dag_info, dag))
            stats.report("colors", tracer_obj.resolve_colors(dag, colors))
            stats.report("shadows", tracer_obj.resolve_shadows(fr.camera, dag_info, dag, shadow_bias, fog_density))
        stats.next_frame()
    return stats
