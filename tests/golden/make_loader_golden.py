"""Generates tests/golden/loader_golden.json by EXECUTING the reference's data loaders
(/root/reference/trainers/loadBinaryMovieLens.py: movieLensData :8-39, gfData :41-62) on small files written here.
The SMB client and the share-path helper are stubbed to the local file system (the loaders only use them to open the
file); everything else -- pandas parsing options, the dropped first row, string ids, pd.unique vocabularies -- is the
reference's own code.  Runs only in the build container; the JSON is committed.   python tests/golden/make_loader_golden.py
"""
import importlib.util
import json
import os
import sys
import tempfile
import types

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "loader_golden.json")

GF = "CUSTOMER_ID,MATERIAL\n0012,00A7\n0013,00A7\n0012,0B01\n0099,00A7\n0013,0B01\n0012,00A7\n"
GF_RZ = ("CUSTOMER_ID,NORMALIZED_CUSTOMER_ID,MATERIAL,PRODUCT_ID,RATING_TYPE\n0012,1,00A7,5,1\n0013,2,00A8,6,0\n"
         "0012,1,00A8,6,1\n0077,3,00A7,5,0\n")
ML = "196\t242\t3\t881250949\n186\t302\t3\t891717742\n22\t377\t1\t878887116\n196\t302\t5\t881250950\n244\t51\t2\t880606923\n"


def load_module():
    for name in ("smbclient", "src", "src.AAUfilename"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["src.AAUfilename"].getAAUfilename = lambda p: p
    sys.modules["smbclient"].open_file = lambda path, mode="r", username=None, password=None: open(path, mode)
    spec = importlib.util.spec_from_file_location("ref_loaders", os.path.join(REF, "trainers/loadBinaryMovieLens.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    L = load_module()
    out = {"inputs": {"gf": GF, "gf_rdzero": GF_RZ, "ml100k": ML}}
    with tempfile.TemporaryDirectory() as tmp:
        for key, text, rd in (("gf", GF, False), ("gf_rdzero", GF_RZ, True)):
            path = os.path.join(tmp, key + ".csv")
            open(path, "w").write(text)
            r = L.gfData(path, "user", "password", rdZero=rd)
            out[key] = {"users": [str(x) for x in r["ratings"]["CUSTOMER_ID"]], "items": [str(x) for x in r["ratings"]["MATERIAL"]],
                        "usersId": [str(x) for x in r["usersId"]], "materialsId": [str(x) for x in r["materialsId"]],
                        "nbrUser": int(r["nbrUser"]), "nbrMaterial": int(r["nbrMaterial"])}
            if rd:
                out[key]["values"] = [float(x) for x in r["ratings"]["RATING_TYPE"]]
        os.makedirs(os.path.join(tmp, "data", "ml-100k"))
        open(os.path.join(tmp, "data", "ml-100k", "u.data"), "w").write(ML)
        cwd = os.getcwd()
        os.chdir(tmp)
        try:
            r = L.movieLensData(1, 0, 0.0)
        finally:
            os.chdir(cwd)
        out["ml100k"] = {"users": [str(x) for x in r["ratings"]["user_id"]], "items": [str(x) for x in r["ratings"]["movie_id"]],
                         "usersId": [str(x) for x in r["usersId"]], "moviesId": [str(x) for x in r["moviesId"]],
                         "nbrUser": int(r["nbrUser"]), "nbrMovie": int(r["nbrMovie"]),
                         "ratings": [float(x) for x in r["ratings"]["rating"]]}
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
