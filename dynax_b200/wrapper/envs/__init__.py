from .rc import RC, EnvStates

__all__ = ["RC", "EnvStates"]
