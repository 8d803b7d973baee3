"""Error types of the likelihood API.

Names and meaning follow the reference's error contract (blueice/exceptions.py:1-32, SURVEY.md
section 8b "Error conventions") so that `except blueice.exceptions.X` code keeps working.
"""


class BlueIceException(Exception):
    """Root of every error raised deliberately by this package."""


def _error(name, doc):
    return type(name, (BlueIceException,), {"__doc__": doc, "__module__": __name__})


NoOpimizationNecessary = _error(
    "NoOpimizationNecessary", "make_objective found no free parameter (reference spelling kept).")
OptimizationFailed = _error(
    "OptimizationFailed", "Both the default minimiser and the Nelder-Mead retry reported failure.")
NotPreparedException = _error(
    "NotPreparedException", "prepare() or set_data() has to be called first.")
NoShapeParameters = _error(
    "NoShapeParameters", "A morpher was constructed without any shape parameter.")
InvalidParameter = _error(
    "InvalidParameter", "A keyword passed to the likelihood is not a known shape or rate parameter.")
InvalidParameterSpecification = _error(
    "InvalidParameterSpecification", "add_shape_parameter / add_rate_parameter was called inconsistently.")
PDFNotComputedException = _error(
    "PDFNotComputedException", "A source was asked for pdf values before its pdf was computed.")
