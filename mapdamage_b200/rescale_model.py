"""Damage-probability table and the exact rescale look-up table.

``rescale._get_corr_prob`` (reference ``rescale.py:23-46``) reads
``Stats_out_MCMC_correct_prob.csv``; ``_rescale_qual_read`` then computes, for
every C->T / G->A column, a new Phred score from the old one and the
position-specific damage probability (``rescale.py:228-246``).  The new score
is a pure function of ``(type, position, Q)``, so it is tabulated here on the
host *with the reference's own Python expressions* -- the device only indexes
the table, which makes the rescaled qualities bit-exact rather than +-1.
"""
import csv
import math

import numpy as np

MAX_PHRED = 93  # '~' - 33, the top of the SAM quality range
LUT_INVALID = 255


class RescaleError(RuntimeError):
    """Mirror of the reference's ``rescale.RescaleError`` (``rescale.py:9``)."""


def get_corr_prob(filepath, rescale_length_5p, rescale_length_3p):
    """``{(ref_nt, read_nt, position): probability}`` -- ``rescale.py:23-46``."""
    try:
        with open(filepath, newline="") as handle:
            reader = csv.DictReader(handle, strict=True)
            corr_prob = {}
            for line in reader:
                position = int(line["Position"])
                if -rescale_length_3p <= position <= rescale_length_5p:
                    corr_prob[("C", "T", position)] = float(line["C.T"])
                    corr_prob[("G", "A", position)] = float(line["G.A"])
            return corr_prob
    except FileNotFoundError:
        raise RescaleError("File does not exist; please re-run mapDamage")
    except csv.Error as error:
        raise RescaleError("Error while reading line %d: %s" % (reader.line_num, error))


def _rescaled_phred(qual, corr):
    """New Phred score of one base; the arithmetic of ``rescale.py:13-20,231-243``."""
    ch = chr(qual + 33)
    pdam = 1 - corr
    pseq = 1 - 10 ** (-(float(ord(ch)) - float(33)) / 10)
    newp = pdam * pseq
    try:
        return int(round(-10 * math.log10(abs(1 - newp))))
    except ValueError:
        return None


class RescaleModel:
    """Dense, device-ready form of the correction table.

    Slot 0 means "no entry" (probability 0, ``corr_prob.get(..., 0)``); 5' position
    ``p`` in ``1..len5p`` is slot ``p``; 3' position ``-p`` (``p`` in ``1..len3p``)
    is slot ``len5p + p``.  ``lut[type][slot][Q]`` is the new Phred score for
    ``type`` 0 = C->T, 1 = G->A; ``inc[type][slot]`` is what one rescaled base
    adds to the ``MR`` sum (``1 - pdam``, ``rescale.py:244``).
    """

    def __init__(self, corr_prob, rescale_length_5p, rescale_length_3p):
        self.len5p = int(rescale_length_5p)
        self.len3p = int(rescale_length_3p)
        self.corr_prob = dict(corr_prob)
        n_slots = self.n_slots = 1 + self.len5p + self.len3p
        self.lut = np.full((2, n_slots, MAX_PHRED + 1), LUT_INVALID, dtype=np.uint8)
        self.inc = np.zeros((2, n_slots), dtype=np.float64)
        self.prob = np.zeros((2, n_slots), dtype=np.float64)
        for t, key in enumerate((("C", "T"), ("G", "A"))):
            for slot in range(n_slots):
                corr = self.corr_prob.get(key + (self.position_of(slot),), 0) if slot else 0
                self.prob[t, slot] = corr
                self.inc[t, slot] = 1 - (1 - corr)
                for qual in range(MAX_PHRED + 1):
                    value = _rescaled_phred(qual, corr)
                    if value is not None and 0 <= value < LUT_INVALID:
                        self.lut[t, slot, qual] = value

    def position_of(self, slot):
        return slot if slot <= self.len5p else -(slot - self.len5p)

    def slot_of(self, position):
        if 0 < position <= self.len5p:
            return position
        if 0 < -position <= self.len3p:
            return self.len5p - position
        return 0

    @classmethod
    def from_csv(cls, filepath, rescale_length_5p, rescale_length_3p):
        corr = get_corr_prob(filepath, rescale_length_5p, rescale_length_3p)
        return cls(corr, rescale_length_5p, rescale_length_3p)
