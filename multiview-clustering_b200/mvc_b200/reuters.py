"""Reuters-21578 ingestion: the .sgm files -> three sparse count views (body, title, category tags) in CSR form,
ready for Sampler.upload_view_csr.

Restates /root/reference/dataset/reuters/data pre-process.R (which needs R with tm and XML) so that BASELINE
configs[1] / [4] run without R:
  :7-40    documents = the pieces of every reut2-*.sgm between </REUTERS> marks that contain "<REUTERS"; per document
           the text between <TITLE>..</TITLE>, between <BODY>..</BODY> and every <D>..</D> field (all category
           kinds: topics, places, people, orgs, exchanges, companies)
  :49-64   body view: lower-case, drop punctuation, drop digits, drop the tm "en" stop words, squeeze white space;
           terms of length >= 3 present in >= 5 documents
  :67-83   title view: the same with a document-frequency floor of 3
  :86-102  tag view: binary indicator per category string
Like the R script, entities (&lt;) are not decoded and apostrophes vanish with the punctuation before the stop words
are removed (so "don't" survives as "dont").  This is host-side data preparation, not part of the sweep.
"""
from __future__ import annotations

import re
import string
from collections import Counter
from pathlib import Path

import numpy as np

# tm::stopwords("en") (the snowball list, 174 entries)
STOPWORDS_EN = """i me my myself we our ours ourselves you your yours yourself yourselves he him his himself she her hers
herself it its itself they them their theirs themselves what which who whom this that these those am is are was were be
been being have has had having do does did doing would should could ought i'm you're he's she's it's we're they're i've
you've we've they've i'd you'd he'd she'd we'd they'd i'll you'll he'll she'll we'll they'll isn't aren't wasn't weren't
hasn't haven't hadn't doesn't don't didn't won't wouldn't shan't shouldn't can't cannot couldn't mustn't let's that's
who's what's here's there's when's where's why's how's a an the and but if or because as until while of at by for with
about against between into through during before after above below to from up down in out on off over under again
further then once here there when where why how all any both each few more most other some such no nor not only own same
so than too very""".split()

_PUNCT = str.maketrans("", "", string.punctuation)
_DIGITS = str.maketrans("", "", string.digits)
_STOP = set(STOPWORDS_EN)


def parse_sgm(dirpath):
    """[(newid, title or None, body or None, [tags])] in file order (data pre-process.R:7-40)."""
    docs = []
    for f in sorted(Path(dirpath).glob("*.sgm")):
        raw = f.read_bytes().decode("latin-1")
        for piece in raw.split("</REUTERS>"):
            if "<REUTERS" not in piece:
                continue
            m = re.search(r'NEWID="([0-9]+)"', piece)
            t = re.search(r"<TITLE>(.*?)</TITLE>", piece, re.S)
            b = re.search(r"<BODY>(.*?)</BODY>", piece, re.S)
            tags = re.findall(r"<D>(.*?)</D>", piece, re.S)
            docs.append((int(m.group(1)) if m else -1, t.group(1) if t else None, b.group(1) if b else None, tags))
    return docs


def tokens(text):
    """tm pipeline of :52-56 followed by the DocumentTermMatrix tokenizer with wordLengths = c(3, Inf)."""
    if text is None:
        return []
    s = text.lower().translate(_PUNCT).translate(_DIGITS)
    return [w for w in s.split() if w not in _STOP and len(w) >= 3]


def _dtm(token_lists, min_docs):
    df = Counter()
    for toks in token_lists:
        df.update(set(toks))
    vocab = sorted(w for w, c in df.items() if c >= min_docs)
    index = {w: j for j, w in enumerate(vocab)}
    rowptr, col, val = [0], [], []
    for toks in token_lists:
        cnt = Counter(index[w] for w in toks if w in index)
        for j in sorted(cnt):
            col.append(j)
            val.append(cnt[j])
        rowptr.append(len(col))
    return {"rowptr": np.asarray(rowptr, np.int32), "col": np.asarray(col, np.int32), "val": np.asarray(val, np.float32),
            "vocab": len(vocab), "terms": vocab}


def load_reuters(dirpath):
    """{"ids", "body", "title", "tags"}: the three CSR count views of data pre-process.R:104-108."""
    docs = parse_sgm(dirpath)
    body = _dtm([tokens(d[2]) for d in docs], 5)
    title = _dtm([tokens(d[1]) for d in docs], 3)
    tagv = sorted({t for d in docs for t in d[3] if t != ""})
    tindex = {t: j for j, t in enumerate(tagv)}
    rowptr, col = [0], []
    for d in docs:
        col += sorted({tindex[t] for t in d[3] if t != ""})
        rowptr.append(len(col))
    tags = {"rowptr": np.asarray(rowptr, np.int32), "col": np.asarray(col, np.int32),
            "val": np.ones(len(col), np.float32), "vocab": len(tagv), "terms": tagv}
    return {"ids": np.asarray([d[0] for d in docs], np.int64), "body": body, "title": title, "tags": tags}


def synthetic_like_reuters(n=21578, seed=1999, k_true=20):
    """Views with the measured shapes of the real collection (body 21 578 x ~13.0k words, ~47 per
    row; title ~5.7k words, ~5 per row; 445 binary tags, ~2 per row; ~12 % empty bodies) drawn from planted topics,
    for boxes where the .sgm files are not available.  Returns ([body, title, tags], labels)."""
    rng = np.random.default_rng(seed)
    z = rng.integers(0, k_true, n)
    out = []
    for vocab, mean_len, empty, binary in ((13000, 47.0, 0.12, False), (5700, 5.0, 0.03, False), (445, 1.9, 0.05, True)):
        theta = rng.dirichlet(np.full(vocab, 0.02), k_true)
        cdf = np.cumsum(theta, axis=1)
        rowptr, col, val = [0], [], []
        lens = np.where(rng.random(n) < empty, 0, rng.poisson(mean_len, n))
        for i in range(n):
            if lens[i]:
                w = np.searchsorted(cdf[z[i]], rng.random(lens[i]))
                w = np.minimum(w, vocab - 1)
                u, c = np.unique(w, return_counts=True)
                col.append(u)
                val.append(np.ones_like(c) if binary else c)
                rowptr.append(rowptr[-1] + len(u))
            else:
                rowptr.append(rowptr[-1])
        out.append({"rowptr": np.asarray(rowptr, np.int32), "col": np.concatenate(col).astype(np.int32),
                    "val": np.concatenate(val).astype(np.float32), "vocab": vocab})
    return out, z


# ---------------------------------------------------------------------------------------------------------------------
# Cache of the ingested collection.  The .sgm files live under /root/reference and do not travel to a GPU box; the three
# CSR views (9 MB) do: `python -m mvc_b200.reuters <sgm dir> <out.npz>` writes them once (data_cache/ is git-ignored but
# ships with the working tree), load_cached() reads them back where the sampler runs.
# ---------------------------------------------------------------------------------------------------------------------
CACHE = __import__("pathlib").Path(__file__).resolve().parents[2] / "data_cache" / "reuters21578_csr.npz"


def save_cache(r, path=CACHE):
    path = __import__("pathlib").Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    out = {"ids": r["ids"]}
    for name in ("body", "title", "tags"):
        for k in ("rowptr", "col", "val"):
            out[f"{name}_{k}"] = r[name][k]
        out[f"{name}_vocab"] = np.int64(r[name]["vocab"])
    np.savez_compressed(path, **out)
    return path


def load_cached(path=CACHE):
    """The three count views [body, title, tags] as CSR dicts, or None when the cache has not been written."""
    path = __import__("pathlib").Path(path)
    if not path.exists():
        return None
    z = np.load(path)
    views = [{"rowptr": z[f"{n}_rowptr"].astype(np.int32), "col": z[f"{n}_col"].astype(np.int32),
              "val": z[f"{n}_val"].astype(np.float32), "vocab": int(z[f"{n}_vocab"])} for n in ("body", "title", "tags")]
    return views, z["ids"]


if __name__ == "__main__":
    import sys
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/dataset/reuters/reuters21578"
    dst = sys.argv[2] if len(sys.argv) > 2 else CACHE
    print(save_cache(load_reuters(src), dst))
