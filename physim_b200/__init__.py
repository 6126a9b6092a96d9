"""physim_b200 — B200-native gravity hot path of jhb123/physim (astro / astro2 / simple_astro + verlet)."""
from .entity import ACCELERATION, ENTITY, accelerations, entities  # noqa: F401
