"""TEST INFRASTRUCTURE — minimal restatement of the monai==0.7.0 symbols the reference imports.

The reference (GhTara/Dose_Prediction, requirements.txt:5) pins monai==0.7.0, which is not
installed in the build image and cannot be fetched (no network).  This package restates, from the
published 0.7.0 sources, only the classes/functions the hot-path model files import, so that the
reference's own `.py` files can be imported UNMODIFIED in the build container to pin our oracle
(`oracle/torch_ref.py`) and to generate `tests/golden/` fixtures.

Parity status: "parity unpinned" w.r.t. real monai — nothing here could be diffed against the real
package.  Nothing under `oracle/` is imported by the product (`dose_prediction_b200/`).
"""
__version__ = "0.7.0+compat"
