from .problem import Bound, BoundKind, Constraint, ConstraintOp, EllPError, Problem, Variable, VariableId  # noqa: F401
