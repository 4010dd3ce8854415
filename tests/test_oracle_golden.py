"""The CPU oracle against the golden vectors made by the unmodified reference.

These pin the oracle (oracle/mdg_oracle.c) -- and with it the table writers,
the SAM reader, the batch builder and the rescale model -- to what
/root/reference actually produces (tests/golden/, oracle/gen_golden.py).
"""
import numpy as np
import pytest

import oracle
from conftest import golden_cases
from helpers import (assert_tables_equal, expected_rescale, load_counting_case,
                     load_rescale_case, render_tables)
from mapdamage_b200.batch import BAMError
from mapdamage_b200.rescale_model import RescaleModel, get_corr_prob


@pytest.mark.parametrize("case_dir,params", golden_cases("counting"))
def test_counting_matches_reference(case_dir, params, tmp_path):
    if params["exception"]:
        assert params["exception"].startswith("BAMError")
        with pytest.raises(BAMError) as info:
            load_counting_case(case_dir, params, tmp_path)
        assert str(info.value) == params["exception"].split(": ", 1)[1]
        return
    batch, reference, libraries, _ = load_counting_case(case_dir, params, tmp_path)
    L, A = params["length"], params["around"]
    mis, comp, lg = oracle.count(batch, reference, length=L, around=A, minqual=params["minqual"],
                                 n_lib=len(libraries))
    render_tables(tmp_path / "out", libraries, L, A, mis, comp, lg)
    assert_tables_equal(tmp_path / "out", case_dir)


def test_counting_threads_sum(tmp_path, golden_dir):
    import json
    case = golden_dir / "fuzz_0_l70_a10_q0"
    params = json.loads((case / "params.json").read_text())
    batch, reference, libraries, _ = load_counting_case(case, params, tmp_path)
    one = oracle.count(batch, reference, n_lib=len(libraries), threads=1)
    four = oracle.count(batch, reference, n_lib=len(libraries), threads=4)
    for a, b in zip(one, four):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("case_dir,params", golden_cases("rescale"))
def test_rescale_matches_reference(case_dir, params):
    batch, reference, _, records = load_rescale_case(case_dir)
    corr = get_corr_prob(case_dir / "Stats_out_MCMC_correct_prob.csv",
                         params["length_5p"], params["length_3p"])
    if params["exception"]:
        assert "MR" in records[0].tags  # host-level check, see test_host_rescale
        return
    qual, mr, status, subs, rc = oracle.rescale(batch, reference, corr)
    if params["rc"] != 0:
        assert rc == -3
        return
    assert rc == 0
    want = expected_rescale(case_dir)
    assert len(want) == batch.n
    for i, (want_qual, want_mr) in enumerate(want):
        off, n = int(batch.base_off[i]), int(batch.l_seq[i])
        if want_qual is None:
            assert batch.qualities_of(i) is None
        else:
            got = (qual[off:off + n] + 33).astype(np.uint8).tobytes().decode("latin-1")
            assert got == want_qual, "record %d (%s)" % (i, records[i].qname)
        if want_mr is None:
            assert status[i] == 0
        else:
            assert status[i] == 1
            assert mr[i] == want_mr, "record %d MR %r != %r" % (i, mr[i], want_mr)
    n_warn = sum("longer than the actual read" in m for m in params["log"])
    assert subs.n_too_long == n_warn


def test_rescale_lut_is_the_reference_arithmetic():
    """Host LUT (Python expressions) == libm evaluation in the C oracle, every cell."""
    corr = {("C", "T", p): 0.9 * 0.67 ** (p - 1) for p in range(1, 13)}
    corr.update({("G", "A", -p): 0.8 * 0.6 ** (p - 1) for p in range(1, 13)})
    corr[("C", "T", -1)] = 0.02
    model = RescaleModel(corr, 12, 12)
    for t in range(2):
        for slot in range(model.n_slots):
            for q in range(94):
                assert model.lut[t, slot, q] == oracle.rescaled_phred(q, model.prob[t, slot])
    assert model.lut[0, 0, 93] == 93
    assert model.slot_of(0) == 0 and model.slot_of(13) == 0 and model.slot_of(-12) == 24
