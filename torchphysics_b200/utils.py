"""``tp.utils`` of the reference: differential operators + ``UserFunction``."""
from .operators import (convective, div, grad, jac, laplacian, matrix_div, normal_derivative, partial, rot,
                        sym_grad)
from .userfn import DomainUserFunction, UserFunction

__all__ = ["laplacian", "grad", "div", "jac", "partial", "convective", "rot", "normal_derivative", "sym_grad",
           "matrix_div", "UserFunction", "DomainUserFunction"]
