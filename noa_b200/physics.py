"""noa::pms scalar types, constants and elements (reference: src/noa/pms/physics.hh:29-94).

Values are the reference's, digit for digit.  `Material` (a mass-fraction mix of elements) has no
counterpart in the reference, whose API knows single elements only (physics.hh:39-43); it follows
the per-element table mixing of PUMAS (src/noa/3rdparty/_pumas/pumas.c:8054-8078) and exists for
BASELINE.json's water / multi-material configurations.
"""
from typing import NamedTuple, Tuple


class AtomicElement(NamedTuple):
    """physics.hh:39-43 -- A [g/mol], I [GeV] (mean excitation), Z."""
    A: float
    I: float
    Z: int


class Material(NamedTuple):
    name: str
    elements: Tuple[AtomicElement, ...]
    fractions: Tuple[float, ...]  # mass fractions, same order


AVOGADRO_NUMBER = 6.02214076E+23     # physics.hh:54
ATOMIC_MASS_ENERGY = 0.931494        # physics.hh:55
ELECTRON_MASS = 0.510998910E-03      # physics.hh:57, GeV/c^2
MUON_MASS = 0.10565839               # physics.hh:58
TAU_MASS = 1.77682                   # physics.hh:59
MUON_CTAU = 658.654                  # physics.hh:61
TAU_CTAU = 87.03E-06                 # physics.hh:62
LARMOR_FACTOR = 0.299792458          # physics.hh:64

STANDARD_ROCK = AtomicElement(22., 0.1364E-6, 11)   # physics.hh:67-71

X_FRACTION = 5E-02                   # physics.hh:76
NPR = 4                              # physics.hh:86

# Builder-supplied elements (PDG values as used by the PUMAS MDF files);
# NOT defined in the reference.
HYDROGEN = AtomicElement(1.0087, 19.2E-9, 1)
OXYGEN = AtomicElement(15.999, 95.0E-9, 8)
IRON = AtomicElement(55.845, 286E-9, 26)
LEAD = AtomicElement(207.2, 823E-9, 82)

WATER = Material("water", (HYDROGEN, OXYGEN), (0.111894, 0.888106))
ROCK = Material("standard_rock", (STANDARD_ROCK,), (1.0,))
IRON_MATERIAL = Material("iron", (IRON,), (1.0,))
LEAD_MATERIAL = Material("lead", (LEAD,), (1.0,))
SWEEP_MATERIALS = (WATER, ROCK, IRON_MATERIAL, LEAD_MATERIAL)   # BASELINE.json config 5
