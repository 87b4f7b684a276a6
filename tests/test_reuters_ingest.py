"""Reuters ingestion (.sgm -> CSR count views) restating dataset/reuters/data pre-process.R.
The collection lives under /root/reference (this container only): skipped where it is absent."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "multiview-clustering_b200"))
SGM = Path("/root/reference/dataset/reuters/reuters21578")


def test_tokeniser_follows_the_tm_pipeline():
    from mvc_b200 import reuters
    # lower-case, punctuation and digits deleted (not replaced), stop words removed afterwards, length >= 3
    assert reuters.tokens("The U.S. didn't cut 1,200 jobs; it's CO-OP's &lt;ABC&gt; plan.") == \
        ["didnt", "cut", "jobs", "coops", "ltabcgt", "plan"]
    assert reuters.tokens(None) == []
    assert len(reuters.STOPWORDS_EN) == 174


def test_synthetic_views_have_reuters_shapes():
    from mvc_b200 import reuters
    views, z = reuters.synthetic_like_reuters(n=2000, seed=3)
    assert [v["vocab"] for v in views] == [13000, 5700, 445]
    for v in views:
        assert v["rowptr"][0] == 0 and v["rowptr"][-1] == len(v["col"]) == len(v["val"])
        assert (np.diff(v["rowptr"]) >= 0).all() and v["col"].max() < v["vocab"]
    assert 35 < len(views[0]["col"]) / 2000 < 50 and set(np.unique(views[2]["val"])) == {1.0}


@pytest.mark.skipif(not SGM.exists(), reason="Reuters .sgm files not present on this box")
def test_reuters_collection_shapes():
    from mvc_b200 import reuters
    r = reuters.load_reuters(SGM)
    n = len(r["ids"])
    assert n == 21578 and sorted(r["ids"]) == list(range(1, 21579))
    body, title, tags = r["body"], r["title"], r["tags"]
    for v in (body, title, tags):
        assert len(v["rowptr"]) == n + 1 and v["rowptr"][-1] == len(v["col"])
        for i in (0, 1, n - 1):                                    # columns ascending inside a row
            c = v["col"][v["rowptr"][i]:v["rowptr"][i + 1]]
            assert (np.diff(c) > 0).all()
    # SURVEY.md §8d [PROBE] shapes (a Python re-tokenisation of the same rules): body ~13.0k words / ~1.01M nonzeros,
    # title ~5.7k words, 445 category strings
    assert tags["vocab"] == 445 and set(np.unique(tags["val"])) == {1.0}
    assert 12000 < body["vocab"] < 14500 and 0.9e6 < len(body["col"]) < 1.15e6
    assert 5000 < title["vocab"] < 6500 and 0.09e6 < len(title["col"]) < 0.13e6
    empty_bodies = int((np.diff(body["rowptr"]) == 0).sum())
    assert 2300 < empty_bodies < 2800
    # every term kept appears in at least 5 (body) / 3 (title) documents
    assert np.bincount(body["col"], minlength=body["vocab"]).min() >= 5
    assert np.bincount(title["col"], minlength=title["vocab"]).min() >= 3
