"""HBV 1.1p — B200-native drop-in for ``hydrodl2/models/hbv/hbv_1_1p.py:8-608``."""

from ... import _cabi as A
from ._packed import PackedHbv


class Hbv_1_1p(PackedHbv):
    """HBV 1.1p: HBV 1.0 + always-on ET shape parameter ``parBETAET`` and capillary
    rise ``parC`` (hbv_1_1p.py:100-101,472-490), extra ``capillary`` flux."""

    _variant = A.VARIANT_HBV11P
    _name = 'HBV 1.1p'
    _capillary = True
    _always_betaet = True
    _extra_bounds = {'parBETAET': [0.3, 5], 'parC': [0, 1]}
