"""oracle shim (TEST INFRASTRUCTURE ONLY) standing in for pip `compressai`.

The reference imports `compressai._CXX`, `compressai.ans`, `compressai.ops` and
`compressai.{available_entropy_coders,get_entropy_coder}` at run time
(/root/reference/cra5/models/compressai/entropy_models/entropy_models.py:42-43,53,62,84).
A same-lineage source copy is vendored inside the reference; this shim points the
package path at it and at oracle/_ref/compressai (where build_ref.py puts the two
pybind modules compiled from the reference's own cpp_exts sources).
"""
import os

_ref = os.environ.get("CRA5_REFERENCE_ROOT", "/root/reference")
__path__.append(os.path.join(_ref, "cra5", "models", "compressai"))
_built = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))),
                      "_ref", "compressai")
__path__.append(_built)


def available_entropy_coders():
    return ["ans"]


def get_entropy_coder():
    return "ans"
