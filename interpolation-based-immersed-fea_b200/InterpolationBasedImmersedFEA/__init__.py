"""Drop-in replacement package for the reference's ``InterpolationBasedImmersedFEA`` on the extraction
hot path: put ``interpolation-based-immersed-fea_b200/`` ahead of the reference on ``sys.path`` and the
demos' ``from InterpolationBasedImmersedFEA.la_utils import *`` / ``...common import *`` resolve here."""
