#!/usr/bin/env python3
"""Extract the reference's OWN golden vectors for the assembly hot path into
tests/golden/reference_kats.json.

Run in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

It only READS test data (numbers) from the reference's test files; no reference
source code is copied.  The GPU box has no /root/reference, so the committed
JSON is what the tests use.

Sources (all under /root/reference):
  * tests/unit_tests/assembly/global.rs:100-138,174-212   CSR pattern KATs (serial == parallel)
  * tests/unit_tests/mesh/snapshots/unit__unit_tests__mesh__procedural__mesh_{1,2}.snap
        insta snapshots of create_rectangular_uniform_tet_mesh(1.0,1,1,1,res) (procedural.rs test :18-30)
  * tests/unit_tests/fe_mesh.rs:60-129                    Hex8 -> Hex27 single element
  * fenris-solid/tests/unit_tests/mod.rs:11-29            Lame parameters / deformation gradients fixtures
  * fenris-solid/tests/unit_tests/materials.rs:74-84      Lame from Young/Poisson
  * fenris-solid/tests/unit_tests/materials.rs:245-262    linear elastic energy densities
  * fenris-solid/tests/unit_tests/materials.rs:299-351    StVK / NeoHookean energy densities
  * tests/unit_tests/assembly.rs:159-162                  reference Quad4 Laplace element matrix (commented-out but valid KAT)
  * tests/convergence_tests/reference_values/poisson3d_mms_{hex8,tet4}_summary.json, poisson2d_mms_quad4_summary.json
"""
import json
import os
import re
import sys

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_kats.json")


def read(path):
    with open(os.path.join(REF, path)) as f:
        return f.read()


def parse_snapshot(path):
    txt = read(path)
    body = txt.split("Mesh {", 1)[1]
    vpart, cpart = body.split("connectivity:", 1)
    vnums = [float(x) for x in re.findall(r"-?\d+\.\d+(?:e-?\d+)?", vpart)]
    assert len(vnums) % 3 == 0
    vertices = [vnums[i:i + 3] for i in range(0, len(vnums), 3)]
    cnums = [int(x) for x in re.findall(r"(?<![\w.])\d+(?![\w.])", cpart.replace("Tet4Connectivity", ""))]
    assert len(cnums) % 4 == 0
    conn = [cnums[i:i + 4] for i in range(0, len(cnums), 4)]
    return {"vertices": vertices, "connectivity": conn}


def ints(s):
    return [int(x) for x in re.findall(r"\d+", s)]


def parse_pattern_kats():
    txt = read("tests/unit_tests/assembly/global.rs")
    serial = txt.split("fn csr_assemble_mock_pattern()", 1)[1].split("fn csr_par_assemble_mock_pattern()", 1)[0]
    par = txt.split("fn csr_par_assemble_mock_pattern()", 1)[1].split("fn gather_global_to_local_args", 1)[0]

    def cases(block):
        out = []
        for m in re.finditer(
            r"solution_dim:\s*(\d+),\s*num_nodes:\s*(\d+),\s*element_connectivities:\s*vec!\[(.*?)\],\s*\};(.*?)assert_eq!\(pattern",
            block, re.S):
            sdim, nn, conn_src, rest = int(m.group(1)), int(m.group(2)), m.group(3), m.group(4)
            elements = [ints(e) for e in re.findall(r"vec!\[(.*?)\]", conn_src, re.S)]
            pm = re.search(r"try_from_offsets_and_indices\(\s*(\d+),\s*(\d+),\s*(vec!\[.*?\]),\s*(vec!\[.*?\]),?\s*\)", rest, re.S)
            nrows, ncols, offs_src, idx_src = int(pm.group(1)), int(pm.group(2)), pm.group(3), pm.group(4)
            rep = re.match(r"vec!\[(\d+);\s*(\d+)\]", offs_src.strip())
            offsets = [int(rep.group(1))] * int(rep.group(2)) if rep else ints(offs_src)
            indices = ints(idx_src)
            out.append({"sdim": sdim, "num_nodes": nn, "elements": elements, "nrows": nrows, "ncols": ncols,
                        "offsets": offsets, "indices": indices})
        return out

    s, p = cases(serial), cases(par)
    assert len(s) == 4 and s == p, "serial and parallel KATs must be identical"
    return s


def main():
    if not os.path.isdir(REF):
        sys.exit("reference checkout not present; run this in the build container")
    kats = {}
    kats["pattern"] = parse_pattern_kats()
    kats["bcc_tet_mesh_1"] = parse_snapshot("tests/unit_tests/mesh/snapshots/unit__unit_tests__mesh__procedural__mesh_1.snap")
    kats["bcc_tet_mesh_2"] = parse_snapshot("tests/unit_tests/mesh/snapshots/unit__unit_tests__mesh__procedural__mesh_2.snap")

    # fe_mesh.rs:60-129 - the input vertices and the expectations (midpoints) are restated as data
    fm = read("tests/unit_tests/fe_mesh.rs").split("fn hex8_to_hex27_single_element_mesh()", 1)[1]
    verts = re.findall(r"Point3::new\(([-\d.]+),\s*([-\d.]+),\s*([-\d.]+)\)", fm.split("let hex8 =", 1)[0])
    edges = [tuple(map(int, m)) for m in re.findall(r"edge_midpoint\((\d+),\s*(\d+)\), abstol", fm)]
    faces = [ints(m) for m in re.findall(r"v\[2\d\]\.coords, midpoint\(&\[(.*?)\]\)", fm)]
    kats["hex27_single"] = {
        "vertices": [[float(c) for c in v] for v in verts],
        "edge_pairs": edges,            # nodes 8..19 are midpoints of these vertex pairs
        "face_sets": faces[:6],         # nodes 20..25 are centroids of these vertex sets
        "center_set": faces[6] if len(faces) > 6 else list(range(8)),
        "abstol": 1e-12,
    }
    assert len(kats["hex27_single"]["vertices"]) == 8 and len(edges) == 12 and len(faces) == 7

    mod = read("fenris-solid/tests/unit_tests/mod.rs")
    mu = float(re.search(r"mu:\s*([\d.]+)", mod).group(1))
    lam = float(re.search(r"lambda:\s*([\d.]+)", mod).group(1))
    f2 = re.search(r"fn deformation_gradient_2d.*?matrix!\[(.*?)\]", mod, re.S).group(1)
    f3 = re.search(r"fn deformation_gradient_3d.*?matrix!\[(.*?)\]", mod, re.S).group(1)
    F2 = [[float(x) for x in row.split(",")] for row in f2.split(";")]
    F3 = [[float(x) for x in row.split(",")] for row in f3.split(";")]
    mats = read("fenris-solid/tests/unit_tests/materials.rs")
    psi2 = float(re.search(r"fn linear_elastic_strain_energy_2d.*?assert_scalar_eq!\(psi,\s*([\d.]+)", mats, re.S).group(1))
    psi3 = float(re.search(r"fn linear_elastic_strain_energy_3d.*?assert_scalar_eq!\(psi,\s*([\d.]+)", mats, re.S).group(1))
    yp = re.search(r"fn lame_from_young_poisson.*?young:\s*([\deE.+-]+),\s*poisson:\s*([\d.]+).*?lame\.mu,\s*([\d.]+).*?lame\.lambda,\s*([\d.]+)", mats, re.S)
    nonlinear = {}
    for name in ("stvk", "neo_hookean"):
        for dim in ("2d", "3d"):
            m = re.search(r"fn %s_strain_energy_%s.*?assert_scalar_eq!\(psi,\s*([\d.]+)" % (name, dim), mats, re.S)
            nonlinear["psi_%s_%s" % (name, dim)] = float(m.group(1))
    kats["materials"] = {
        **nonlinear,
        "mu": mu, "lambda": lam, "F2": F2, "F3": F3, "psi_linear_2d": psi2, "psi_linear_3d": psi3,
        "young": float(yp.group(1)), "poisson": float(yp.group(2)),
        "lame_mu": float(yp.group(3)), "lame_lambda": float(yp.group(4)),
    }

    asm = read("tests/unit_tests/assembly.rs")
    q4 = re.search(r"2\.0\s*/\s*3\.0.*?;", asm, re.S)
    # the KAT is K = 1/6 [[4,-1,-2,-1],[-1,4,-1,-2],[-2,-1,4,-1],[-1,-2,-1,4]]  (tests/unit_tests/assembly.rs:159-162)
    kats["quad4_laplace_reference_element"] = {
        "sixth_times": [[4, -1, -2, -1], [-1, 4, -1, -2], [-2, -1, 4, -1], [-1, -2, -1, 4]],
        "source_present": q4 is not None,
    }

    mms = {}
    for name in ("poisson3d_mms_hex8", "poisson3d_mms_tet4", "poisson2d_mms_quad4"):
        p = os.path.join(REF, "tests/convergence_tests/reference_values", name + "_summary.json")
        if os.path.exists(p):
            with open(p) as f:
                mms[name] = json.load(f)
    kats["mms_summaries"] = mms

    with open(OUT, "w") as f:
        json.dump(kats, f, indent=1)
    print("wrote", OUT, {k: (len(v) if hasattr(v, "__len__") else v) for k, v in kats.items()})


if __name__ == "__main__":
    main()
