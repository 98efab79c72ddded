"""CPU-only checks of the product's host side: the C-ABI library loads and exports every symbol the header
declares, the host-side generators/tables restate the reference exactly, and the product refuses to run
without a GPU (no CPU fallback).  No compute calls are made here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import fenris_b200 as fb
from fenris_b200 import _native as nat
from oracle import fenris_oracle as fo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        c = fb.Context(0)
        c.close()
        return True
    except fb.Fb200Error:
        return False


def test_library_exports_every_declared_symbol():
    lib = nat.lib()
    declared = nat.declared_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/fenris_b200.h but not exported"
    assert lib.fb200_abi_version() == 1
    assert lib.fb200_status_string(0) == b"ok"


def test_header_cites_reference_interfaces():
    src = open(os.path.join(ROOT, "include", "fenris_b200.h")).read()
    for cite in ("global.rs:65", "global.rs:133", "314-376", "elliptic.rs", "coloring.rs:6-70", "quadrature_table.rs"):
        assert cite in src


def test_no_cpu_fallback_without_gpu():
    if _has_gpu():
        pytest.skip("a GPU is present")
    with pytest.raises(fb.Fb200Error) as ei:
        fb.Context(0)
    assert ei.value.status == fb.ERR_CUDA
    mesh = fb.create_unit_box_uniform_hex_mesh_3d(2)
    with pytest.raises(fb.Fb200Error):
        fb.CsrAssembler()
    with pytest.raises(fb.Fb200Error):
        fb.color_nodes(mesh)


def test_integration_binding_lists_every_entry_point():
    """INTEGRATION.md's `extern "C"` block (the Rust -sys crate a maintainer would add) names exactly the functions the header declares."""
    txt = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blk = txt[txt.index('extern "C" {'):txt.index("`usize == u64`")]
    bound = set(re.findall(r"pub fn (fb200_[a-z0-9_]+)", blk))
    assert bound == set(nat.declared_symbols())
    hdr = open(os.path.join(ROOT, "include", "fenris_b200.h")).read()
    for name, value in re.findall(r"\b(FB200_(?:ERR_[A-Z_]+|OK))\s*=\s*(\d+)", hdr):
        assert re.search(rf"pub const {name}: i32 = {value};", txt), name


def test_header_is_plain_c_and_the_c_example_links(tmp_path):
    """The boundary is a C ABI: include/fenris_b200.h must compile as C99 (what cgo / bindgen / JNI headers consume), and the plain C
    program examples/assemble_csr.c must link against the library with nothing but gcc.  Without a device it has to fail loudly at
    fb200_create (exit code 3) - there is no CPU path behind the ABI."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    exe = str(tmp_path / "assemble_csr")
    libdir = os.path.join(ROOT, "fenris_b200", "lib")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "assemble_csr.c"), "-L", libdir, "-lfenris_b200", f"-Wl,-rpath,{libdir}", "-lm", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")  # also on a GPU box: this test is about the failure path
    r = subprocess.run([exe, "2"], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 3, (r.returncode, r.stdout, r.stderr)
    assert "fb200_create" in r.stderr and "CUDA" in r.stderr


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "fenris_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, re.M), f
                assert "cpu_ref" not in txt and "fenris_oracle" not in txt, f


@pytest.mark.parametrize("n", [1, 2, 5])
def test_generators_restate_reference_exactly(kats, n):
    m = fb.create_unit_box_uniform_tet_mesh_3d(n)
    v, c = fo.create_unit_box_uniform_tet_mesh_3d(n)
    assert np.array_equal(m.vertices(), v) and np.array_equal(m.connectivity().astype(np.int64), c)
    # Tet10Mesh::from(&tet4) (mesh_convert.rs:42-83): library converter against the literal restatement - same labels, same midpoints
    m10 = fb.tet10_mesh_from(m)
    v10, c10 = fo.tet10_mesh_from_tet4(v, c)
    assert np.array_equal(m10.connectivity().astype(np.int64), c10) and np.array_equal(m10.vertices(), v10)
    if n <= 2:  # the reference's insta snapshots
        snap = kats[f"bcc_tet_mesh_{n}"]
        assert m.vertices().tolist() == snap["vertices"] and m.connectivity().tolist() == snap["connectivity"]
    m = fb.create_unit_box_uniform_hex_mesh_3d(n)
    v, c = fo.create_unit_box_uniform_hex_mesh_3d(n)
    assert np.array_equal(m.vertices(), v) and np.array_equal(m.connectivity().astype(np.int64), c)
    m27 = fb.hex27_mesh_from(m)
    v27, c27 = fo.hex27_mesh_from_hex8(v, c)
    assert np.array_equal(m27.connectivity().astype(np.int64), c27) and np.allclose(m27.vertices(), v27, atol=1e-15, rtol=0)
    m = fb.create_unit_square_uniform_quad_mesh_2d(n)
    v, c = fo.create_unit_square_uniform_quad_mesh_2d(n)
    assert np.array_equal(m.vertices(), v) and np.array_equal(m.connectivity().astype(np.int64), c)


def test_rectangular_generators_and_empty():
    m = fb.create_rectangular_uniform_hex_mesh(2.0, 1, 2, 3, 2)
    v, c = fo.create_rectangular_uniform_hex_mesh(2.0, 1, 2, 3, 2)
    assert np.array_equal(m.vertices(), v) and np.array_equal(m.connectivity().astype(np.int64), c)
    m = fb.create_rectangular_uniform_tet_mesh(1.5, 2, 1, 3, 2)
    v, c = fo.create_rectangular_uniform_tet_mesh(1.5, 2, 1, 3, 2)
    assert np.array_equal(m.vertices(), v) and np.array_equal(m.connectivity().astype(np.int64), c)
    assert fb.create_unit_box_uniform_hex_mesh_3d(0).num_elements() == 0
    assert fb.create_unit_box_uniform_tet_mesh_3d(0).num_elements() == 0


def test_hex27_single_element_kat(kats):
    # tests/unit_tests/fe_mesh.rs:60-129
    k = kats["hex27_single"]
    V = np.array(k["vertices"])
    m27 = fb.hex27_mesh_from(fb.Mesh(V, np.arange(8, dtype=np.uint64)[None, :], fb.HEX8))
    assert m27.connectivity()[0].tolist() == list(range(27))
    v = m27.vertices()
    for i, (a, b) in enumerate(k["edge_pairs"]):
        assert np.allclose(v[8 + i], (V[a] + V[b]) / 2, atol=k["abstol"], rtol=0)
    for i, fs in enumerate(k["face_sets"]):
        assert np.allclose(v[20 + i], V[fs].mean(axis=0), atol=k["abstol"], rtol=0)
    assert np.allclose(v[26], V.mean(axis=0), atol=k["abstol"], rtol=0)


def test_canonical_quadrature_and_lame(kats):
    for et in (fb.QUAD4, fb.TET4, fb.HEX8, fb.HEX27, fb.TET10):
        w, p = fb.canonical_stiffness_quadrature(et)
        wo, po = fo.canonical_stiffness_rule(et)
        assert np.array_equal(w, wo) and np.array_equal(p, po)
    m = kats["materials"]
    lame = fb.LameParameters.from_young_poisson(fb.YoungPoisson(m["young"], m["poisson"]))
    assert abs(lame.mu - m["lame_mu"]) <= 4 * np.spacing(m["lame_mu"])
    assert abs(lame.lambda_ - m["lame_lambda"]) <= 4 * np.spacing(m["lame_lambda"])


def test_builder_and_assembler_surface():
    mesh = fb.create_unit_box_uniform_hex_mesh_3d(2)
    qt = fb.UniformQuadratureTable.from_quadrature(fb.canonical_stiffness_quadrature(fb.HEX8))
    ea = (fb.ElementEllipticAssemblerBuilder().with_finite_element_space(mesh).with_operator(fb.LaplaceOperator())
          .with_quadrature_table(qt).with_u(np.zeros(mesh.num_nodes())).build())
    assert ea.solution_dim() == 1 and ea.num_elements() == 8 and ea.num_nodes() == 27 and ea.element_node_count(0) == 8
    buf = np.zeros(8, dtype=np.uint64)
    ea.populate_element_nodes(buf, 0)
    assert buf.tolist() == [0, 1, 4, 3, 9, 10, 13, 12]
    lame = fb.LameParameters(1.0, 2.0)
    el = (fb.ElementEllipticAssemblerBuilder().with_finite_element_space(mesh)
          .with_operator(fb.MaterialEllipticOperator(fb.LinearElasticMaterial())).with_quadrature_table(qt.with_uniform_data(lame))
          .with_u(np.zeros(3 * mesh.num_nodes())).build())
    assert el.solution_dim() == 3 and el._data().shape == (8, 2)
    with pytest.raises(AssertionError):  # "Local element dofs (u_element) dimension mismatch", elliptic.rs:378-383
        fb.ElementEllipticAssemblerBuilder().with_finite_element_space(mesh).with_operator(fb.LaplaceOperator()) \
            .with_quadrature_table(qt).with_u(np.zeros(5)).build()
    with pytest.raises(fb.Fb200Error):  # non-linear materials have no device specialisation and there is no CPU fallback
        fb.MaterialEllipticOperator(object())
    with pytest.raises(AssertionError):
        fb.UniformQuadratureTable.from_points_and_weights(np.zeros((3, 3)), np.zeros(2))


def test_host_helpers_size_queries():
    L = nat.lib()
    nv, ne = C.c_uint64(0), C.c_uint64(0)
    assert L.fb200_gen_tet_mesh(44, 44, 44, C.c_double(1 / 44), C.byref(nv), C.byref(ne), None, None) == 0
    assert (nv.value, ne.value) == (176309, 1022208)  # BASELINE config C2
    assert L.fb200_gen_hex_mesh(126, 126, 126, C.c_double(1 / 126), C.byref(nv), C.byref(ne), None, None) == 0
    assert (nv.value, ne.value) == (2048383, 2000376)  # BASELINE config C3
    assert L.fb200_gen_tet_mesh(161, 161, 161, C.c_double(1 / 161), C.byref(nv), C.byref(ne), None, None) == 0
    assert (nv.value, ne.value) == (8424809, 50079372)  # BASELINE config C5
