"""Data loaders -- drop-in mirror of the reference's trainers/loadBinaryMovieLens.py (movieLensData :8-39, gfData
:41-62).  Same names, arguments and result dicts; "ratings" is a dict of columns (what the reference turns its frame
into anyway: `dict(testData)`, trainers/twoTower.py:186-187) instead of a DataFrame, and the SMB share
(`smbc.open_file(getAAUfilename(...))`) becomes the local file system -- username / psw are accepted and ignored.
Parsing goes through interactions.read_csv_columns (schemas pinned against the executed reference loaders,
tests/golden/loader_golden.json); vocabularies are in order of first appearance, like `pd.unique`.
"""
import numpy as np

from .interactions import read_csv_columns


def _unique(column):
    return np.array(list(dict.fromkeys(np.asarray(column).tolist())), dtype=object)


def gfData(filename, username=None, psw=None, rdZero=False):
    """{"ratings": {"CUSTOMER_ID", "MATERIAL"[, "RATING_TYPE"]}, "nbrUser", "nbrMaterial", "materialsId", "usersId"}
    of a Grundfos file: string ids, first row dropped (:49-50), RATING_TYPE as float with rdZero (:51-53)."""
    cols = read_csv_columns(filename, "twotower-rdzero" if rdZero else "twotower")
    ratings = {"CUSTOMER_ID": [str(x) for x in cols["user"]], "MATERIAL": [str(x) for x in cols["item"]]}
    if rdZero:
        ratings["RATING_TYPE"] = [float(x) for x in cols["value"]]
    materialId, usersId = _unique(ratings["MATERIAL"]), _unique(ratings["CUSTOMER_ID"])
    return {"ratings": ratings, "nbrUser": len(usersId), "nbrMaterial": len(materialId), "materialsId": materialId,
            "usersId": usersId}


def movieLensData(ratedVal, unratedVal, zeroProb, path='data/ml-100k/u.data'):
    """ml-100k u.data (:8-21): ids as strings, every rating replaced by float(ratedVal); realRat = the set of
    (user, movie) pairs (:23-25).  unratedVal / zeroProb belong to the block the reference has commented out (:27-37)."""
    cols = read_csv_columns(path, "ml-100k")
    users, movies = [str(x) for x in cols["user"]], [str(x) for x in cols["item"]]
    ratings = {"movie_id": movies, "user_id": users, "rating": [float(ratedVal)] * len(users)}
    moviesId, usersId = _unique(movies), _unique(users)
    return {"ratings": ratings, "nbrUser": len(usersId), "nbrMovie": len(moviesId), "realRat": set(zip(users, movies)),
            "moviesId": moviesId, "usersId": usersId}
