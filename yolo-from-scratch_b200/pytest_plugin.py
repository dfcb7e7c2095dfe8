"""pytest plugin: `pytest -p yolo_from_scratch_b200.pytest_plugin <reference>/tests` runs the reference's own,
unchanged test-suite with the B200 path swapped in.  The import hook must be active before the
test modules (and the reference's conftest.py, tests/conftest.py:16) import `train`."""
from .install import enable_import_hook

enable_import_hook()
