"""Host-side mesh / DOF interface of the batched Cook's-membrane path.

Mirrors the slice of the reference's ``fem_preprocess.PreProcessing`` that the
hot path and its callers read (upstream src/fem_preprocess.py:24-30 global
dicts, :114-289 ``get_input_data``, :291-443 ``assign_dof_parfor_topopt``,
:445-509 ``assign_storage_topopt``; cards from model_property_cards.py:5-75):
same class-level dictionaries, same key names, same 1-based index arrays, so
``fem_test.py``-style scripts and ``fem_postprocess`` read results where they
expect them.  Only Q4 / 2x2 Gauss / plane strain / linear elasticity is
covered -- the single configuration the reference's hot path ever runs.
"""
from __future__ import annotations

import io
import os

import numpy as np


def default_cards():
    """The constants of model_property_cards.py:25-57 that the hot path reads."""
    material = [{"id": 1, "type": 1, "E": 20.0, "v": 0.3, "hsv": 0, "in_flag": 0}]
    section = [{"id": 1, "type": 1, "int_scheme": "guass", "intp": 2, "thk": 10, "etype": 1, "stype": 2,
                "eform": 1, "estorage": 0, "nalpha": 0, "mode_type": 0}]
    part = [{"id": 1, "sec_id": 1, "mat_id": 1, "body": np.zeros((3, 1))}]
    solution_control = {
        "solver": 1, "large_disp_flag": 0, "print_flag": 0,
        "load_control": {"numsteps": 1},
        "nr_param": {"max_iter": 10, "tol_cr": 1.0e-10, "tol_Rforce": 0},
    }
    return material, section, part, solution_control, "2D"


def _records(lines, start, conv, ncol=None):
    rows = []
    k = start
    while k < len(lines) and lines[k].strip():
        tok = lines[k].split()
        try:
            rows.append([conv(t) for t in (tok if ncol is None else tok[:ncol])])
        except ValueError:
            break
        k += 1
    return rows


def parse_feap(text: str):
    """Parse the reference's FEAP-style input (mixed LF/CRLF): header on line 2
    (nnodes nele nmat ndm ndf nen), then the 'COORdinates ALL', 'ELEMents ALL',
    'BOUNdary conditions' and 'FORCe conditions' blocks.  The generation flag in
    column 2 of every record is discarded, as upstream does
    (src/fem_preprocess.py:213-221)."""
    lines = text.replace("\r\n", "\n").replace("\r", "\n").split("\n")
    head = lines[1].split()
    nnodes, nele, _nmat, ndm, ndf, nen = (int(t) for t in head[:6])
    if (ndm, ndf, nen) != (2, 2, 4):
        raise ValueError("only 2-D four-node quadrilateral meshes with 2 dofs per node are supported")
    where = {}
    for i, ln in enumerate(lines):
        key = ln.strip()[:4].upper()
        if key in ("COOR", "ELEM", "BOUN", "FORC") and key not in where:
            where[key] = i
    for key in ("COOR", "ELEM"):
        if key not in where:
            raise ValueError(f"mesh file has no {key} block")
    xyz = np.asarray(_records(lines, where["COOR"] + 1, float, 4), dtype=np.float64)
    con = np.asarray(_records(lines, where["ELEM"] + 1, int, 3 + nen), dtype=np.int64)
    if xyz.shape[0] != nnodes or con.shape[0] != nele:
        raise ValueError("node/element count does not match the header")
    coord = xyz[:, [0, 2, 3]]                 # (id, x, y)
    ien = con[:, 3:3 + nen]                   # (id, gen, mat, n1..n4) -> nodes
    bc = np.asarray(_records(lines, where["BOUN"] + 1, int, 4), dtype=np.int64) if "BOUN" in where else None
    ld = np.asarray(_records(lines, where["FORC"] + 1, float, 4), dtype=np.float64) if "FORC" in where else None
    support = bc[:, [0, 2, 3]] if bc is not None and bc.size else np.zeros((0, 3), np.int64)
    load = ld[:, [0, 2, 3]] if ld is not None and ld.size else np.zeros((0, 3))
    return {"nnodes": nnodes, "nele": nele, "coord": coord, "IEN": ien, "support": support, "nodal_load": load}


def cook_membrane_feap(nx: int, ny: int, total_load: float = 50.0) -> str:
    """Cook's membrane with nx x ny elements in the reference's input format:
    corners (0,0),(48,44),(48,60),(0,44); nodes row-major with nx+1 per row;
    the x=0 edge clamped; ``total_load`` in +y lumped on the x=48 edge.
    nx=20, ny=10 reproduces Armero_cooksm_20x10.txt (7 significant digits)."""
    out = io.StringIO()
    out.write("FEAP * * PLANE strain problem\n")
    out.write("%10d%10d%10d%10d%10d%10d\n \n\n" % ((nx + 1) * (ny + 1), nx * ny, 1, 2, 2, 4))
    out.write("COORdinates ALL\n")
    for j in range(ny + 1):
        for i in range(nx + 1):
            xi, eta = i / nx, j / ny
            x = 48.0 * xi
            y = 44.0 * xi + eta * (44.0 * (1.0 - xi) + 16.0 * xi)
            out.write("%9d 0 % .6E % .6E\n" % (j * (nx + 1) + i + 1, x, y))
    out.write("\nELEMents ALL\n")
    for j in range(ny):
        for i in range(nx):
            n1 = j * (nx + 1) + i + 1
            out.write("%6d   0     1 %6d %6d %6d %6d\n" % (j * nx + i + 1, n1, n1 + 1, n1 + nx + 2, n1 + nx + 1))
    out.write("\nBOUNdary conditions\n")
    for j in range(ny + 1):
        out.write("%10d   0   1   1\n" % (j * (nx + 1) + 1))
    out.write("\nFORCe conditions\n")
    for j in range(ny + 1):
        w = total_load / ny * (0.5 if j in (0, ny) else 1.0)
        out.write("%9d 0 % .6E % .6E\n" % (j * (nx + 1) + nx + 1, 0.0, w))
    out.write(" \nEND\n")
    return out.getvalue()


class PreProcessing:
    """Class-level state shared by solver, post-processing and callers
    (upstream src/fem_preprocess.py:24-30)."""

    model_data: dict = {}
    out_data: dict = {}
    sol_data: dict = {}
    topo_data: dict = {}

    # src/fem_preprocess.py:32-42
    _t, _o = 0.666666666666667, 0.333333333333333
    Pdevs = np.zeros((9, 9))
    for _r, _c, _v in [(0, 0, _t), (0, 4, -_o), (0, 8, -_o), (4, 0, -_o), (4, 4, _t), (4, 8, -_o), (8, 0, -_o),
                       (8, 4, -_o), (8, 8, _t), (1, 1, .5), (1, 3, .5), (3, 1, .5), (3, 3, .5), (2, 2, .5),
                       (2, 6, .5), (6, 2, .5), (6, 6, .5), (5, 5, .5), (5, 7, .5), (7, 5, .5), (7, 7, .5)]:
        Pdevs[_r, _c] = _v
    del _r, _c, _v

    @classmethod
    def reset(cls):
        cls.model_data, cls.out_data, cls.sol_data, cls.topo_data = {}, {}, {}, {}

    @classmethod
    def modeldata_initialization_topopt(cls, infile_name, model_file_name=None):
        """Same entry point as upstream (src/fem_preprocess.py:56-112): parse the
        mesh, attach the cards, number the dofs, allocate result storage.
        ``infile_name`` may be a path or the file's text."""
        if isinstance(infile_name, str) and ("\n" in infile_name):
            text = infile_name
        else:
            with open(os.fspath(infile_name), "r", newline="") as f:
                text = f.read()
        cls.reset()
        cls.get_input_data(text)
        md = cls.model_data
        if not md["loading"]["nodal_load"][:, 1:].any():
            raise ValueError("There is neither applied displacement nor load.")
        md["material"], md["section"], md["part"], md["solution_control"], md["ele_type"] = default_cards()
        cls.assign_dof_parfor_topopt()
        cls.assign_storage_topopt()
        if model_file_name:
            import scipy.io as sio
            flat = {"coord": md["mesh_info"]["coord"]}
            flat.update({k: v for k, v in md["dof_info"].items() if isinstance(v, (np.ndarray, int))})
            flat["Pf"] = md["loading"]["Pf"]
            sio.savemat(model_file_name, {"model_data": flat})
        return md

    @classmethod
    def get_input_data(cls, text):
        m = parse_feap(text)
        cls.model_data = {
            "mesh_info": {"nnodes": m["nnodes"], "nele": m["nele"], "coord": m["coord"], "max_node_dof": 2,
                          "max_ele_node": 4},
            "element": [{"id": e + 1, "nnodes": 4, "nodes": m["IEN"][e].copy(), "part_id": 1}
                        for e in range(m["nele"])],
            "support": m["support"],
            "loading": {"nodal_load": m["nodal_load"], "nodal_disp": np.zeros((0, 3))},
        }

    @classmethod
    def assign_dof_parfor_topopt(cls):
        """DOF maps of src/fem_preprocess.py:291-443: ID[c, n] = 2 n + c + 1,
        LM[:, e] = ID[:, IEN[e]] flattened column-major, free/supp sets, Pf."""
        md = cls.model_data
        nn, ne = md["mesh_info"]["nnodes"], md["mesh_info"]["nele"]
        ndof = 2 * nn
        ID = (2 * np.arange(nn)[None, :] + np.arange(2)[:, None] + 1).astype(np.int64)
        IEN = np.stack([el["nodes"] for el in md["element"]]).astype(np.int64)
        LM = np.stack([ID[:, IEN[e] - 1].flatten(order="F") for e in range(ne)], axis=1)
        fixed = np.zeros(ndof, dtype=bool)
        for n, fx, fy in md["support"]:
            if fx == 1:
                fixed[ID[0, n - 1] - 1] = True
            if fy == 1:
                fixed[ID[1, n - 1] - 1] = True
        P = np.zeros(ndof)
        for n, px, py in md["loading"]["nodal_load"]:
            n = int(n)
            if px != 0:
                P[ID[0, n - 1] - 1] += px
            if py != 0:
                P[ID[1, n - 1] - 1] += py
        all_dof = np.arange(1, ndof + 1, dtype=np.int64)
        free_dof, supp_dof = all_dof[~fixed], all_dof[fixed]
        md["dof_info"] = {"LM": LM, "ID": ID, "IEN": IEN, "all_dof": all_dof, "free_dof": free_dof,
                          "supp_dof": supp_dof, "ndof": ndof, "nsupp": int(supp_dof.size),
                          "nfree": int(free_dof.size)}
        md["loading"]["Pf"] = P[free_dof - 1].reshape(-1, 1)
        md["loading"]["Ps"] = P[supp_dof - 1].reshape(-1, 1)
        md["loading"]["Us"] = np.zeros((supp_dof.size, 1))

    @classmethod
    def assign_storage_topopt(cls):
        """Result containers with the upstream layout (src/fem_preprocess.py:445-509):
        ele_stress/ele_strain[6, nip, nele, numsteps+1], out_data['step'] list."""
        md = cls.model_data
        nn, ne = md["mesh_info"]["nnodes"], md["mesh_info"]["nele"]
        nsteps = md["solution_control"]["load_control"]["numsteps"]
        od = cls.out_data
        od["ele_stress"] = np.zeros((6, 4, ne, nsteps + 1))
        od["ele_strain"] = np.zeros((6, 4, ne, nsteps + 1))
        od["step"] = [{"nodal_disp": np.zeros((3, nn)), "nodal_force": np.zeros((3, nn)),
                       "Uf": np.zeros((md["dof_info"]["nfree"], 1)), "Us": np.zeros((md["dof_info"]["nsupp"], 1)),
                       "Pf": np.zeros((md["dof_info"]["nfree"], 1)), "Ps": np.zeros((md["dof_info"]["nsupp"], 1))}]
