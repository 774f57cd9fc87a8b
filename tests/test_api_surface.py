"""Drop-in surface: every public function, class and method of the reference files on the hot path and beside it
(tests/golden/api_surface.json, read from the reference's syntax trees by tests/golden/make_api_surface.py) exists under
the same name in the mirror modules of binary-recommendation_b200/.  CPU only: names, not behaviour -- behaviour is
what the parity tests hold."""
import importlib
import json
import os

import pytest

SURFACE = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "api_surface.json")))

# Names of the reference that are deliberately absent, each with its reason.
ABSENT = {}


@pytest.mark.parametrize("path", sorted(SURFACE))
def test_mirror_has_every_reference_name(path):
    entry = SURFACE[path]
    mod = importlib.import_module("binrec_b200." + entry["mirror"])
    missing = [f for f in entry["functions"] if not hasattr(mod, f)]
    for cls_name, methods in entry["classes"].items():
        cls = getattr(mod, cls_name, None)
        if cls is None:
            missing.append("class " + cls_name)
            continue
        missing += [f"{cls_name}.{m}" for m in methods if not hasattr(cls, m)]
    missing = [m for m in missing if (path, m) not in ABSENT]
    assert not missing, f"{path} -> binrec_b200.{entry['mirror']} lacks {missing}"


def test_surface_covers_every_file_of_the_models_and_trainers_directories():
    """The reference's src/models/ and trainers/ hold exactly these files."""
    files = {p for p in SURFACE if p.startswith(("src/models/", "trainers/"))}
    assert files == {"src/models/RModel.py", "src/models/NeuMFModel.py", "src/models/BPRModel.py", "src/models/NCFModel.py",
                     "src/models/bpr.py", "trainers/twoTower.py", "trainers/topKmetrics.py", "trainers/loadBinaryMovieLens.py"}
