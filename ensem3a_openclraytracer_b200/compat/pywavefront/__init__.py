"""Stand-in for `pywavefront` on hosts that do not have it.

The reference's scene importer (FileManager.py:260-261,297-304) uses pywavefront for exactly one thing:
the raw `v`, `vn` and `vt` records of the OBJ file in file order (`.vertices`, `.parser.normals`,
`.parser.tex_coords`; `.materials` is fetched and never read) — faces and materials are parsed by
FileManager itself.  This module provides that and nothing else; it is used only when put on PYTHONPATH
explicitly."""


class _Parser(object):
    def __init__(self):
        self.normals = []
        self.tex_coords = []


class Wavefront(object):

    def __init__(self, file_name, strict=False, encoding="utf-8", create_materials=False, collect_faces=False,
                 parse=True, cache=False):
        self.file_name = file_name
        self.vertices = []
        self.parser = _Parser()
        self.materials = {}
        self.meshes = {}
        self.mesh_list = []
        if parse:
            self.parse()

    def parse(self):
        with open(self.file_name) as fh:
            for line in fh:
                tok = line.split()
                if not tok:
                    continue
                if tok[0] == "v":
                    self.vertices.append(tuple(float(t) for t in tok[1:4]))
                elif tok[0] == "vn":
                    self.parser.normals.append(tuple(float(t) for t in tok[1:4]))
                elif tok[0] == "vt":
                    self.parser.tex_coords.append(tuple(float(t) for t in tok[1:3]))
