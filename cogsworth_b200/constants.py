"""Unit constants of the galactic-dynamics hot path.

gala's ``galactic`` unit system is (kpc, Myr, Msun, rad); cogsworth feeds positions in
kpc and velocities in km/s (reference ``src/cogsworth/pop.py:1221-1226``) and builds the
kicked-orbit time grid in seconds (``src/cogsworth/kicks.py:118``).  Every float below is
the IEEE double astropy/gala would produce for the same conversion.
"""

# 1 Julian year = 365.25 d = 31 557 600 s; both products are exact in FP64.
SEC_PER_MYR = 3.15576e13
SEC_PER_GYR = 3.15576e16
MYR_PER_GYR = 1000.0

# astropy converts s -> Myr by multiplying with the rounded reciprocal
# (``UnitBase._to``: self.scale / other.scale with self.scale == 1.0).
MYR_PER_SEC = 1.0 / SEC_PER_MYR
# Myr -> Gyr scale used when ``t1[Gyr] + tphys*u.Myr`` is formed at kicks.py:134
GYR_PER_MYR = SEC_PER_MYR / SEC_PER_GYR

# 1 kpc/Myr in km/s  (kpc = 3.0856775814913673e19 m)
KMS_PER_KPC_MYR = 977.7922216807891
KPC_MYR_PER_KMS = 1.0 / KMS_PER_KPC_MYR

# G in kpc^3 Msun^-1 Myr^-2 (CODATA-2018 G = 6.6743e-11, IAU nominal solar mass parameter)
G_GALACTIC = 4.498502151469553e-12
