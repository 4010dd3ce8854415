"""Down-sampling of the kept reads (``reader.py:84-96,134-164``): the native continuation of CPython's
``random.Random`` stream against the interpreter itself, and -- with the CPU oracle doing the counting -- against
tables the unmodified reference produced with ``-n X --downsample-seed S`` (oracle/gen_golden_downsample.py)."""
import random

import numpy as np
import pytest

import oracle
from conftest import golden_cases
from helpers import assert_tables_equal, load_counting_case, render_tables
from mapdamage_b200 import downsample


def reservoir_by_the_book(count, seed, n, first=0, slots=None):
    """``BAMReader._downsample_to_fixed_number`` (reader.py:144-164) over stream indices."""
    rand = random.Random(seed)
    sample = [None] * count if slots is None else slots
    for index in range(first, first + n):
        if index >= count:
            index, at = rand.randint(0, index), index
            if index >= count:
                continue
            sample[index] = at
        else:
            sample[index] = index
    return sorted(x for x in sample if x is not None)


@pytest.mark.parametrize("seed", [0, 1, 7, 2 ** 40 + 3, -5, None])
def test_fraction_draws_are_cpythons(seed):
    if seed is None:
        seed = random.SystemRandom().getrandbits(64)
    sampler = downsample.FractionSampler(0.37, seed)
    got = np.concatenate([sampler.mask(1000), sampler.mask(0), sampler.mask(1), sampler.mask(5000)])
    rand = random.Random(seed)
    want = np.array([rand.random() < 0.37 for _ in range(6001)])
    assert np.array_equal(got, want)


@pytest.mark.parametrize("count,n", [(10, 5000), (1, 100), (300, 200), (7, 100_000), (64, 64), (5, 0)])
@pytest.mark.parametrize("seed", [0, 5, 2 ** 33])
def test_reservoir_draws_are_cpythons(count, n, seed):
    sampler = downsample.ReservoirSampler(count, seed)
    sampler.feed(n // 3)
    sampler.feed(n - n // 3)
    assert list(sampler.selected()) == reservoir_by_the_book(count, seed, n)


def test_reservoir_beyond_32_bit_indices():
    """``randint(0, index)`` takes two generator words per try once ``index + 1`` needs more than 32 bits, and tries
    again while the value is above ``index``: after the same reads the generator must be where CPython's is."""
    first, n = 2 ** 35 + 11, 3000
    sampler = downsample.ReservoirSampler(8, 3)
    sampler._seen = first
    sampler.feed(n)
    rand = random.Random(3)
    for index in range(first, first + n):
        rand.randint(0, index)
    assert tuple(int(x) for x in sampler._state) == rand.getstate()[1]
    assert not len(sampler.selected())  # 8 slots out of 2^35: no read lands in them


@pytest.mark.parametrize("value", [1.0, -0.1, 2])
def test_fraction_out_of_range(value):
    with pytest.raises(ValueError):
        downsample.FractionSampler(value)


def test_sampler_for_follows_the_option_rule():
    assert downsample.sampler_for(None) is None
    assert isinstance(downsample.sampler_for(0.5, 1), downsample.FractionSampler)
    assert isinstance(downsample.sampler_for(1, 1), downsample.ReservoirSampler)
    assert downsample.sampler_for(12.7, 1).count == 12  # config.py:399-400
    with pytest.raises(ValueError):
        downsample.sampler_for(0)


def test_selection_hands_out_masks_batch_by_batch():
    selection = downsample.Selection([0, 3, 4, 9])
    got = np.concatenate([selection.mask(4), selection.mask(3), selection.mask(0), selection.mask(5)])
    assert list(np.flatnonzero(got)) == [0, 3, 4, 9]


@pytest.mark.parametrize("case_dir,params", golden_cases("counting_downsample"))
def test_sampled_counting_matches_reference(case_dir, params, tmp_path):
    batch, reference, libraries, _ = load_counting_case(case_dir.parent / params["source"], params, tmp_path)
    sampler = downsample.sampler_for(params["downsample"], params["downsample_seed"])
    if isinstance(sampler, downsample.ReservoirSampler):
        sampler.feed(batch.n)
        sampler = downsample.Selection(sampler.selected())
    half = batch.n // 2
    keep = np.concatenate([sampler.mask(half), sampler.mask(batch.n - half)])
    downsample.apply_mask(batch, np.ones(batch.n, dtype=np.bool_), keep)
    L, A = params["length"], params["around"]
    mis, comp, lg = oracle.count(batch, reference, length=L, around=A, minqual=params["minqual"], n_lib=len(libraries))
    render_tables(tmp_path / "out", libraries, L, A, mis, comp, lg)
    assert_tables_equal(tmp_path / "out", case_dir)
