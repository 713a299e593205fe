"""
Error types raised on the solve path.  Same class names, constructor keywords and `details` keys as
the reference's `min_ratio_cycle/exceptions.py` (SolverError :12, ValidationError :68, GraphError :107,
AlgorithmError :161, ResourceExhaustionError :275, TimeoutError :305, NumericalError :343,
NumericalInstabilityError :372, ConvergenceError :397, CycleValidationError :439), so that callers'
`except` clauses and `e.details[...]` lookups keep working.  Only what the hot path and its callers
touch is provided; the reference's ErrorHandler / formatting helpers are out of scope.
"""

from __future__ import annotations

from typing import Any


def _drop_none(**kw) -> dict[str, Any]:
    return {k: v for k, v in kw.items() if v is not None}


class SolverError(Exception):
    """Base class; `details` carries machine-readable context."""

    def __init__(self, message: str, details: dict[str, Any] | None = None, component: str | None = None,
                 suggested_fix: str | None = None, recovery_hint: str | None = None):
        super().__init__(message)
        self.message = message
        self.details = dict(details) if details else {}
        for key, val in (("component", component), ("suggested_fix", suggested_fix),
                         ("recovery_hint", recovery_hint)):
            if val:
                self.details.setdefault(key, val)

    def __str__(self) -> str:
        if not self.details:
            return self.message
        return f"{self.message} ({', '.join(f'{k}={v}' for k, v in self.details.items())})"

    def to_dict(self) -> dict[str, Any]:
        return {"error_type": type(self).__name__, "message": self.message, "details": self.details}


class ValidationError(SolverError):
    def __init__(self, message: str, invalid_value: Any = None, expected_type: type | None = None,
                 valid_range: tuple | None = None):
        details: dict[str, Any] = {}
        if invalid_value is not None:
            details.update(invalid_value=invalid_value, invalid_type=type(invalid_value).__name__)
        if expected_type is not None:
            details["expected_type"] = expected_type.__name__
        if valid_range is not None:
            details["valid_range"] = valid_range
        super().__init__(message, details)
        self.invalid_value, self.expected_type, self.valid_range = invalid_value, expected_type, valid_range


class GraphError(SolverError):
    def __init__(self, message: str, graph_properties: dict[str, Any] | None = None,
                 suggested_fix: str | None = None):
        details = dict(graph_properties or {})
        if suggested_fix:
            details["suggested_fix"] = suggested_fix
        super().__init__(message, details)
        self.graph_properties, self.suggested_fix = graph_properties, suggested_fix


class GraphStructureError(GraphError):
    pass


class AlgorithmError(SolverError):
    def __init__(self, message: str, algorithm_name: str | None = None, iterations: int | None = None,
                 convergence_info: dict[str, Any] | None = None):
        details = _drop_none(algorithm=algorithm_name, iterations=iterations)
        details.update(convergence_info or {})
        super().__init__(message, details)
        self.algorithm_name, self.iterations, self.convergence_info = algorithm_name, iterations, convergence_info


class ConfigurationError(SolverError):
    pass


class ResourceExhaustionError(SolverError):
    def __init__(self, message: str, resource: str, limit: float | None = None, usage: float | None = None,
                 suggested_fix: str | None = None, recovery_hint: str | None = None):
        super().__init__(message, details={"resource": resource, **_drop_none(limit=limit, usage=usage)},
                         suggested_fix=suggested_fix, recovery_hint=recovery_hint)
        self.resource, self.limit, self.usage = resource, limit, usage


class TimeoutError(SolverError):  # noqa: A001 - name fixed by the reference API
    def __init__(self, message: str, time_elapsed: float | None = None, time_limit: float | None = None,
                 operation: str | None = None):
        super().__init__(message, _drop_none(time_elapsed_s=time_elapsed, time_limit_s=time_limit,
                                             operation=operation or None))
        self.time_elapsed, self.time_limit, self.operation = time_elapsed, time_limit, operation


class NumericalError(SolverError):
    def __init__(self, message: str, computation_details: dict[str, Any] | None = None,
                 suggested_mode: str | None = None, **kwargs: Any):
        details = dict(computation_details or {})
        if suggested_mode:
            details["suggested_mode"] = suggested_mode
        super().__init__(message, details, **kwargs)
        self.computation_details, self.suggested_mode = computation_details, suggested_mode


class NumericalInstabilityError(NumericalError):
    def __init__(self, message: str, computation_details: dict[str, Any] | None = None,
                 component: str | None = None, suggested_fix: str | None = None,
                 recovery_hint: str | None = None):
        super().__init__(message, computation_details=computation_details, component=component)
        if suggested_fix:
            self.details.setdefault("suggested_fix", suggested_fix)
        if recovery_hint:
            self.details.setdefault("recovery_hint", recovery_hint)


class ConvergenceError(AlgorithmError):
    def __init__(self, message: str, algorithm_name: str, max_iterations: int, tolerance: float | None = None,
                 final_error: float | None = None):
        super().__init__(message, algorithm_name=algorithm_name, iterations=max_iterations,
                         convergence_info=_drop_none(max_iterations=max_iterations, tolerance=tolerance,
                                                     final_error=final_error))


class CycleValidationError(ValidationError):
    def __init__(self, message: str, cycle: list | None = None, issue: str | None = None):
        super().__init__(message, invalid_value=cycle)
        if issue:
            self.details["issue"] = issue
        self.cycle, self.issue = cycle, issue
