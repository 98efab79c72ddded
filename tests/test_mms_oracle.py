"""End-to-end pin of the ORACLE's global Laplace matrices against the reference's stored convergence results
(tests/convergence_tests/reference_values/poisson{2d,3d}_mms_{quad4,hex8,tet4}_summary.json; 1 % tolerance as in
poisson_mms_common.rs:40-65).  CPU only."""
import pytest

from oracle import fenris_oracle as fo
from tests import mms


@pytest.mark.parametrize("name", ["quad4", "hex8", "tet4"])
def test_oracle_reproduces_reference_mms_errors(kats, name):
    et, producer, qrule, erule, key, resolutions = mms.CASES[name]
    golden = kats["mms_summaries"][key]
    for k, res in enumerate(resolutions):
        v, c = producer(res)
        prob = fo.Problem(et, v, c, fo.LAPLACE, *qrule())
        ro, ci, vals = fo.assemble_fast(prob)
        l2, h1 = mms.solve_poisson(et, v, c, mms.csr_from(ro, ci, vals), qrule(), erule())
        assert abs(l2 - golden["L2_errors"][k]) / golden["L2_errors"][k] < 0.01, (name, res, l2, golden["L2_errors"][k])
        assert abs(h1 - golden["H1_seminorm_errors"][k]) / golden["H1_seminorm_errors"][k] < 0.01, (name, res, h1)
