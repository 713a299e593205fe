"""
B200-native implementation of the parametric negative-cycle hot path of
DiogoRibeiro7/min_ratio_cycle behind the reference's own solver API.
Exported names match the reference package (min_ratio_cycle/__init__.py:7-9).
"""

from .solver import Edge, MinRatioCycleSolver

__all__ = ["MinRatioCycleSolver", "Edge"]
