from .tabular import Tabular

__all__ = ["Tabular"]
