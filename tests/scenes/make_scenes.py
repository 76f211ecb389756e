#!/usr/bin/env python
"""Generates the synthetic stand-in scenes (tutorial scenes are not in the reference repo).
c1_example.xml is configs_example/scene.xml with its constants turned into <default> parameters; the others
are derived from it. Run once; outputs are committed."""
import math
import re

import numpy as np


def mat(m):
    return ' '.join('%.9g' % v for v in np.asarray(m, np.float64).reshape(16))


def T(x, y, z):
    m = np.eye(4); m[:3, 3] = [x, y, z]; return m


def S(x, y, z):
    return np.diag([x, y, z, 1.0])


def R(axis, deg):
    a = np.asarray(axis, float); a /= np.linalg.norm(a); x, y, z = a
    c, s = math.cos(math.radians(deg)), math.sin(math.radians(deg))
    m = np.eye(4)
    m[:3, :3] = [[c + x * x * (1 - c), x * y * (1 - c) - z * s, x * z * (1 - c) + y * s],
                 [y * x * (1 - c) + z * s, c + y * y * (1 - c), y * z * (1 - c) - x * s],
                 [z * x * (1 - c) - y * s, z * y * (1 - c) + x * s, c + z * z * (1 - c)]]
    return m


def animated_cube(name, bsdf, k0, k1):
    return f'''	<shape type="cube" id="{name}">
		<ref id="{bsdf}" />
		<animation name="to_world">
			<transform time="0">
				<matrix value="{mat(k0)}" />
			</transform>
			<transform time="$T">
				<matrix value="{mat(k1)}" />
			</transform>
		</animation>
	</shape>
'''


def static_cube(name, bsdf, m):
    return f'''	<shape type="cube" id="{name}">
		<ref id="{bsdf}" />
		<transform name="to_world">
			<matrix value="{mat(m)}" />
		</transform>
	</shape>
'''


base = open('c1_example.xml').read()
head = base[:base.index('\t<shape type="cube" id="ShortBox">')]
tail = base[base.index('\t<emitter type="point">'):]

# C3: room + one cube rotating about y by 2 degrees over T (matrix lerp, not slerp)
k0 = T(0.1, 0.45, 0.1) @ R([0, 1, 0], 20) @ S(0.3, 0.45, 0.3)
k1 = T(0.1, 0.45, 0.1) @ R([0, 1, 0], 22) @ S(0.3, 0.45, 0.3)
open('c3_rotor.xml', 'w').write(head + animated_cube('Rotor', 'TallBoxBSDF', k0, k1) + tail)

# C4: room + 32 dominoes, each rotating about its bottom edge, staggered
shapes = []
n = 32
for i in range(n):
    x = -0.93 + 1.86 * i / (n - 1); z = 0.35 * math.sin(i * 0.7)
    hw, hh, hd = 0.018, 0.16, 0.07
    a0 = -2.0 * i * 0.9; a1 = a0 - 3.0

    def key(a):
        return T(x + hw, 0, z) @ R([0, 0, 1], a) @ T(-hw, hh, 0) @ S(hw, hh, hd)
    bsdf = 'ShortBoxBSDF' if i % 3 else ('LeftWallBSDF' if i % 2 else 'RightWallBSDF')
    shapes.append(animated_cube(f'Domino{i}', bsdf, key(a0), key(a1)))
open('c4_domino.xml', 'w').write(head + ''.join(shapes) + tail)

# C5 (small): room whose walls are thin cube slabs (triangle meshes only, no analytic rectangles),
# two moving cubes and a flat-shaded PLY mesh without normals/uvs
pre = base[:base.index('\t<shape type="rectangle" id="Floor">')]
th = 0.02
slabs = [
    static_cube('FloorSlab', 'FloorBSDF', T(0, -th, 0) @ S(1, th, 1)),
    static_cube('CeilSlab', 'CeilingBSDF', T(0, 2 + th, 0) @ S(1, th, 1)),
    static_cube('BackSlab', 'BackWallBSDF', T(0, 1, -1 - th) @ S(1, 1, th)),
    static_cube('RightSlab', 'RightWallBSDF', T(1 + th, 1, 0) @ S(th, 1, 1)),
    static_cube('LeftSlab', 'LeftWallBSDF', T(-1 - th, 1, 0) @ S(th, 1, 1)),
]
movers = base[base.index('\t<shape type="cube" id="ShortBox">'):base.index('\t<emitter type="point">')]
ply = '''	<shape type="ply" id="Gem">
		<string name="filename" value="gem.ply" />
		<boolean name="face_normals" value="true" />
		<ref id="RightWallBSDF" />
		<transform name="to_world">
			<scale value="0.22" />
			<rotate y="1" angle="25" />
			<translate x="0.05" y="1.15" z="0.3" />
		</transform>
	</shape>
'''
open('c5_slabroom.xml', 'w').write(pre + ''.join(slabs) + movers + ply + tail)

# gem.ply: an octahedron with a twisted top, no normals / uvs
v = [(1, 0, 0), (0, 0, 1), (-1, 0, 0), (0, 0, -1), (0.2, 1.3, 0.1), (0, -1, 0)]
f = [(0, 4, 1), (1, 4, 2), (2, 4, 3), (3, 4, 0), (1, 5, 0), (2, 5, 1), (3, 5, 2), (0, 5, 3)]
with open('gem.ply', 'w') as fh:
    fh.write('ply\nformat ascii 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n'
             'element face %d\nproperty list uchar int vertex_indices\nend_header\n' % (len(v), len(f)))
    for p in v:
        fh.write('%g %g %g\n' % p)
    for t in f:
        fh.write('3 %d %d %d\n' % t)

# gem.serialized + c6_serialized.xml: the gem of c5_slabroom.xml as sub-mesh 1 of a version-4, double-precision
# `.serialized` file (sub-mesh 0 is a decoy triangle; normals, uvs and colours are present so that the loader has to
# skip / narrow them). tests/golden/lanes_c6_serialized.json is the reference's own SerializedMesh loader on this file.
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
from mitsuba3dopplertof_b200 import meshio   # noqa: E402
_pos, _faces, _, _ = meshio.load_ply('gem.ply')
_rng = np.random.default_rng(5)
meshio.write_serialized('gem.serialized', [
    {"name": "decoy", "positions": [[0, 0, 0], [1, 0, 0], [0, 1, 0]], "faces": [[0, 1, 2]]},
    {"name": "gem", "positions": _pos, "faces": _faces, "normals": _pos / np.linalg.norm(_pos, axis=1, keepdims=True),
     "texcoords": _rng.random((len(_pos), 2)), "colors": _rng.random((len(_pos), 3))}], version=4, double_precision=True)
with open('c5_slabroom.xml') as fh:
    _xml = fh.read().replace('<shape type="ply" id="Gem">\n\t\t<string name="filename" value="gem.ply" />',
                             '<shape type="serialized" id="Gem">\n\t\t<string name="filename" value="gem.serialized" />\n'
                             '\t\t<integer name="shape_index" value="1" />')
with open('c6_serialized.xml', 'w') as fh:
    fh.write(_xml)
