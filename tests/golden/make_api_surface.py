"""Generates tests/golden/api_surface.json: the public names (module-level functions, classes and their methods) of the
reference files on the hot path and beside it, read from their syntax trees (nothing is executed).  tests/test_api_surface.py
holds the mirrors in binary-recommendation_b200/ to this list, with every omission named and justified there.
Runs only in the build container (the reference tree is not on the GPU box).   python tests/golden/make_api_surface.py
"""
import ast
import json
import os

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "api_surface.json")
FILES = {   # reference file -> mirror module
    "src/models/RModel.py": "RModel",
    "src/models/NeuMFModel.py": "NeuMFModel",
    "src/models/BPRModel.py": "BPRModel",
    "src/models/NCFModel.py": "NCFModel",
    "src/models/bpr.py": "BPRModel",
    "trainers/twoTower.py": "twoTower",
    "trainers/topKmetrics.py": "topKmetrics",
    "trainers/loadBinaryMovieLens.py": "loadBinaryMovieLens",
    "src/origin_models/svd/SVD.py": "SVD",
    "src/origin_models/svd/topKMetrics.py": "topKmetrics",
}


def surface(path):
    tree = ast.parse(open(os.path.join(REF, path)).read())
    out = {"functions": [], "classes": {}}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef):
            out["functions"].append(node.name)
        elif isinstance(node, ast.ClassDef):
            out["classes"][node.name] = [n.name for n in node.body if isinstance(n, ast.FunctionDef)]
    return out


def main():
    data = {path: {"mirror": mod, **surface(path)} for path, mod in FILES.items()}
    with open(OUT, "w") as f:
        json.dump(data, f, indent=1)
    print("wrote", OUT, {p: (len(v["functions"]), {c: len(m) for c, m in v["classes"].items()}) for p, v in data.items()})


if __name__ == "__main__":
    main()
