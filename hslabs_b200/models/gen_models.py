#!/usr/bin/env python3
"""Generate the three legged-model inputs (myant / hexapod / spider) and the
gait preset table as minimal MuJoCo-style XML / text.

Only what the gait-evaluation path reads is emitted (reference
model.cpp:246-289, visualization.cpp:442-504): ``body@pos``, the FIRST
``geom`` of a body (``type``, ``size``, ``fromto`` | ``pos``) and the ``joint``
(``type``, ``pos``, ``axis``).  Everything the reference ignores (rgba, range,
density, actuators, defaults ...) is left out.  The body trees below restate
SURVEY.md Appendix A; tests/test_models.py checks, when /root/reference is
mounted, that these files yield bit-identical model constants to the
reference's own XML files.

Run:  python hslabs_b200/models/gen_models.py     (writes next to this file)
"""
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def _fmt(v):
    return " ".join(repr(float(x)) if x != int(x) else str(int(x)) for x in v)


def _capsule(size, fromto, kind="capsule"):
    return '<geom type="%s" size="%s" fromto="%s"/>' % (kind, repr(size), _fmt(fromto))


def _hinge(axis):
    return '<joint type="hinge" pos="0 0 0" axis="%s"/>' % _fmt(axis)


def _limb(ind, name, top_pos, top_axis, segs):
    """3-hinge limb: segs = [(capsule end point, next body pos)] x3, axes y/z,x,x."""
    out = []
    axes = [top_axis, (1, 0, 0), (1, 0, 0)]
    pos = top_pos
    for d in range(3):
        out.append("%s<body name=\"%s_%d\" pos=\"%s\">" % (ind + "  " * d, name, d, _fmt(pos)))
        out.append("%s  %s" % (ind + "  " * d, _hinge(axes[d])))
        out.append("%s  %s" % (ind + "  " * d, _capsule(0.08, (0, 0, 0) + tuple(segs[d]))))
        pos = segs[d]
    for d in (2, 1, 0):
        out.append("%s</body>" % (ind + "  " * d))
    return out


def _wrap(name, torso_lines):
    return "\n".join(['<mujoco model="%s">' % name, "  <worldbody>"] + torso_lines +
                     ["  </worldbody>", "</mujoco>", ""])


def hexapod():
    L = ['    <body name="torso" pos="0 0 0.9">',
         "      " + _capsule(0.1, (-0.8, 0, 0, 0.8, 0, 0)),
         '      <joint type="free" pos="0 0 0"/>']
    for row, x in (("front", 0.8), ("mid", 0), ("back", -0.8)):
        L.append('      <body name="%s_legs" pos="%s">' % (row, _fmt((x, 0, 0))))
        L.append("        " + _capsule(0.08, (0, -0.2, 0, 0, 0.2, 0)))
        for side, s in (("left", 1), ("right", -1)):
            L += _limb("        ", "%s_%s" % (row, side), (0, 0.2 * s, 0), (0, 1, 0),
                       [(0, 0.05 * s, 0), (0, 0.4 * s, 0), (0, 0.4 * s, 0)])
        L.append("      </body>")
    L.append("    </body>")
    return _wrap("hexapod", L)


def myant():
    L = ['    <body name="torso" pos="0 0 0.9">',
         '      <geom type="sphere" size="0.25" pos="0 0 0"/>',
         '      <joint type="free" pos="0 0 0"/>']
    for name, sx, sy in (("front_left", 1, 1), ("front_right", -1, 1),
                         ("left_back", -1, -1), ("right_back", 1, -1)):
        L.append('      <body name="%s_leg" pos="0 0 0">' % name)
        L.append("        " + _capsule(0.08, (0, 0, 0, 0.4 * sx, 0.2 * sy, 0)))
        L += _limb("        ", name, (0.4 * sx, 0.2 * sy, 0), (0, 1, 0),
                   [(0, 0.05 * sy, 0), (0, 0.4 * sy, 0), (0, 0.4 * sy, 0)])
        L.append("      </body>")
    L.append("    </body>")
    return _wrap("myant", L)


def spider():
    L = ['    <body name="torso" pos="0 0 0.5">',
         "      " + _capsule(0.5, (0, 0, 0, 0, 0, 0.08), kind="cylinder"),
         '      <joint type="free" pos="0 0 0"/>']
    for name, x, y in (("front_left", 0.433, 0.25), ("front_right", 0.433, -0.25),
                       ("mid_left", 0, 0.5), ("mid_right", 0, -0.5),
                       ("back_left", -0.433, 0.25), ("back_right", -0.433, -0.25)):
        s = 1 if y > 0 else -1
        L += _limb("      ", name, (x, y, 0), (0, 0, 1),
                   [(0, 0, -0.1), (0, 0.4 * s, 0), (0, 0.4 * s, 0)])
    L.append("    </body>")
    return _wrap("spider", L)


# Gait presets: (id, model, torso z offset, yaw, step_duration, period,
# step_length, step_height, extras).  Values restate the reference's
# pgs_config.txt rows 0-27 (rows 28-32 name weaver*.xml models that are not in
# the reference tree).  Emitted in the key/value format that
# modelplayer::get_pgs_config_params parses (player.cpp:170-208).
PRESETS = [
    (0, "myant", -.07, 0, .5, .5, .5, .1, ""), (1, "myant", -.07, 0, 1, .5, .5, .1, ""),
    (2, "myant", -.07, 0, 1, .5, .2, .03, ""), (3, "hexapod", -.47, 0, 1, .5, .2, .02, ""),
    (4, "hexapod", -.47, 0, 1, .5, 1e-6, 1e-6, ""), (5, "myant", -.07, 0, 1, .5, 1e-6, 1e-6, ""),
    (6, "hexapod", -.47, 0, 1, 3, .2, .02, ""), (7, "hexapod", -.1, 0, 1, 3, .2, .02, ""),
    (8, "hexapod", -.1, 0, 1, 3, .5, .1, ""), (9, "myant", -.07, 0, 1, 3, .5, .1, ""),
    (10, "hexapod", -.1, 1.571, 1, 3, .5, .1, ""), (11, "hexapod", -.1, 0, 1, .5, 1e-6, 1e-6, ""),
    (12, "hexapod", -.47, 0, 1, 3, .5, .1, ""), (13, "hexapod", -.1, 0, 1, 3, 1e-6, .2, ""),
    (14, "myant", -.07, 0, 1, 3, 1e-6, .1, ""), (15, "myant", -.07, 0, 1, 3, .3, .05, ""),
    (16, "hexapod", -.05, 1.571, 1, 3, .3, .1, ""), (17, "myant", -.07, 1.571, 1, 3, .5, .1, ""),
    (18, "hexapod", -.1, 1.571, 1, 3, .5, .2, ""), (19, "myant", -.07, 1.571, 1, 3, .5, .2, ""),
    (20, "hexapod", -.1, 0, 0, 3, .5, .1, ""), (21, "myant", -.07, 1.571, 1, 3, .15, .1, ""),
    (22, "myant", -.07, 1.571, 1, 3, .5, .05, ""),
    (23, "hexapod", -.1, 0, 1, 3, .5, .1, "curvature 0.5 lateral_foot_shift 0.2"),
    (24, "spider", .1, 0, 1, 3, .5, .1, "curvature -0.15 lateral_foot_shift 0.4"),
    (25, "spider", 0, 1.571, 1, 3, .5, .1, "lateral_foot_shift 0.4"),
    (26, "spider", .05, 0, 1, 3, .5, .1, "curvature -0.05 radial_foot_shift 0.4"),
    (27, "hexapod", -.1, 0, 1, 3, .5, .1, "curvature -0.03"),
]


def presets_text():
    rows = []
    for pid, model, z, yaw, sd, T, Ls, h, extra in PRESETS:
        row = ("%d xml_file %s.xml torso_pos 0 0 %r torso_angles 0 0 %r step_duration %r "
               "period %r step_length %r step_height %r" % (pid, model, z, yaw, sd, T, Ls, h))
        rows.append(row + (" " + extra if extra else ""))
    return "\n".join(rows) + "\n"


def main():
    for name, fn in (("hexapod", hexapod), ("myant", myant), ("spider", spider)):
        with open(os.path.join(HERE, name + ".xml"), "w") as f:
            f.write(fn())
    with open(os.path.join(HERE, "pgs_presets.txt"), "w") as f:
        f.write(presets_text())


if __name__ == "__main__":
    main()
