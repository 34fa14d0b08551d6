"""HBV 1.0 — B200-native drop-in for ``hydrodl2/models/hbv/hbv.py:8-596``."""

from ... import _cabi as A
from ._packed import PackedHbv


class Hbv(PackedHbv):
    """HBV 1.0: 5 states, 12 parameters (+ parBETAET when listed dynamic,
    hbv.py:124-125), daily step, nmul components, gamma-UH routing."""

    _variant = A.VARIANT_HBV
    _name = 'HBV 1.0'
    _capillary = False
    _always_betaet = False
