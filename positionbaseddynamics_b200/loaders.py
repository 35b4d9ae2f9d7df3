"""Mesh file readers with the reference's names and file semantics -- SURVEY.md section 8 row f-4 ("mesh / scene import"), the part the
hot path's callers need: Utilities::TetGenLoader (Utils/TetGenLoader.cpp:10-262: *.tet, TetGen *.node + *.ele, Gmsh-style *.msh) and
Utilities::OBJLoader (Utils/OBJLoader.h:30-155: triangulated OBJ, v / vt / vn / f with 1-based indices, per-axis scale applied to the
positions).  Host-side IO: plain Python + numpy, no engine involved.  Return conventions follow pyPBD (pyPBD/UtilitiesModule.cpp:126-215):
TetGenLoader.* -> (vertices [n, 3] float32, tets flat uint32); OBJLoader.loadObj -> (x, normals, texCoords, faces)."""
import numpy as np


class MeshFaceIndices:
    """Utilities::MeshFaceIndices (Utils/OBJLoader.h:12-18)."""
    def __init__(self, pos, tex=(-1, -1, -1), nor=(-1, -1, -1)):
        self.posIndices, self.texIndices, self.normalIndices = list(pos), list(tex), list(nor)


class VertexData:
    """What pyPBD's loadObjToMesh returns first: the vertex positions (PBD::VertexData subset)."""
    def __init__(self, x): self._x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, 3)
    def size(self): return len(self._x)
    def getVertices(self): return self._x
    def getPosition(self, i): return self._x[i].copy()
    def __array__(self, dtype=None, copy=None): return self._x if dtype is None else self._x.astype(dtype)
    def __len__(self): return len(self._x)


class FaceMesh:
    """The face list of an IndexedFaceMesh as loadObjToMesh fills it (numFaces / getFaces: what addTriangleModel / addRigidBody read)."""
    def __init__(self, faces, n_points): self._f = np.ascontiguousarray(faces, dtype=np.uint32).reshape(-1, 3); self._n = int(n_points)
    def numFaces(self): return len(self._f)
    def numVertices(self): return self._n
    def getFaces(self): return self._f.reshape(-1)


def _lines(filename):
    with open(filename) as f:
        return f.read().splitlines()


class TetGenLoader:
    @staticmethod
    def loadTetFile(filename):
        """*.tet, version 1.2 layout (TetGenLoader.cpp:10-110): five header lines (version, num_materials, num_vertices, num_tetras, num_triangles),
        a MATERIALS label + the materials, a VERTICES label, the vertices, a TETRAS label, the tets as "i0 i1 i2 i3 material" (indices as stored)."""
        ln = _lines(filename)
        nm, nv, nt = int(ln[1].split()[1]), int(ln[2].split()[1]), int(ln[3].split()[1])
        at = 5 + 1 + nm + 1  # header (5), MATERIALS label, materials, VERTICES label
        x = np.array([[float(v) for v in ln[at + i].split()[:3]] for i in range(nv)], dtype=np.float32).reshape(-1, 3)
        at += nv + 1         # TETRAS label
        t = np.array([[int(v) for v in ln[at + i].split()[:4]] for i in range(nt)], dtype=np.uint32).reshape(-1, 4)
        return x, t.reshape(-1)

    @staticmethod
    def loadTetgenModel(nodeFilename, eleFilename):
        """TetGen *.node ("n 3 0 0", then "index x y z") + *.ele ("m 4 0", then "index a b c d"), indices as stored (TetGenLoader.cpp:113-185)."""
        nl, el = _lines(nodeFilename), _lines(eleFilename)
        nv, nt = int(nl[0].split()[0]), int(el[0].split()[0])
        x = np.array([[float(v) for v in nl[1 + i].split()[1:4]] for i in range(nv)], dtype=np.float32).reshape(-1, 3)
        t = np.array([[int(v) for v in el[1 + i].split()[1:5]] for i in range(nt)], dtype=np.uint32).reshape(-1, 4)
        return x, t.reshape(-1)

    @staticmethod
    def loadMSHModel(mshFilename):
        """*.msh as the reference reads it (TetGenLoader.cpp:187-262): a label line, the vertex count, "index x y z" lines, three lines down
        the tet count, "index a b c d" lines with 1-based indices (stored 0-based)."""
        ln = _lines(mshFilename)
        nv = int(ln[1].split()[0])
        x = np.array([[float(v) for v in ln[2 + i].split()[1:4]] for i in range(nv)], dtype=np.float32).reshape(-1, 3)
        at = 2 + nv + 2
        nt = int(ln[at].split()[0])
        t = np.array([[int(v) - 1 for v in ln[at + 1 + i].split()[1:5]] for i in range(nt)], dtype=np.uint32).reshape(-1, 4)
        return x, t.reshape(-1)


class OBJLoader:
    @staticmethod
    def loadObj(filename, scale=(1.0, 1.0, 1.0)):
        """Triangulated OBJ (OBJLoader.h:30-155).  Which index fields a face carries follows from whether vt / vn lines were seen BEFORE it, as in
        the reference (flags set while reading)."""
        x, normals, tex, faces = [], [], [], []
        vt = vn = False
        s = [float(v) for v in scale]
        for line in _lines(filename):
            tok = line.split()
            if not tok:
                continue
            if tok[0] == "v":
                x.append([float(tok[1 + i]) * s[i] for i in range(3)])
            elif tok[0] == "vt":
                tex.append([float(tok[1]), float(tok[2])]); vt = True
            elif tok[0] == "vn":
                normals.append([float(tok[1 + i]) for i in range(3)]); vn = True
            elif tok[0] == "f":
                pos, ti, ni = [], [-1, -1, -1], [-1, -1, -1]
                for k in range(3):
                    parts = [p for p in tok[1 + k].split("/") if p != ""]  # StringTools::tokenize drops empty tokens
                    pos.append(int(parts[0]) - 1)
                    if vn and vt: ti[k] = int(parts[1]) - 1; ni[k] = int(parts[2]) - 1
                    elif vn: ni[k] = int(parts[1]) - 1
                    elif vt: ti[k] = int(parts[1]) - 1
                faces.append(MeshFaceIndices(pos, ti, ni))
        return (np.array(x, dtype=np.float32).reshape(-1, 3), np.array(normals, dtype=np.float32).reshape(-1, 3),
                np.array(tex, dtype=np.float32).reshape(-1, 2), faces)

    @staticmethod
    def loadObjToMesh(filename, scale=(1.0, 1.0, 1.0)):
        """pyPBD's loadObjToMesh (UtilitiesModule.cpp:140-183): (VertexData, face mesh) ready for SimulationModel.addTriangleModel / addRigidBody."""
        x, _, _, faces = OBJLoader.loadObj(filename, scale)
        f = np.array([fc.posIndices for fc in faces], dtype=np.uint32).reshape(-1, 3)
        return VertexData(x), FaceMesh(f, len(x))
