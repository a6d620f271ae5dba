"""CPU test (no GPU): the documents do not cite files or tests that do not exist (round 1 shipped two dangling test
citations).  Every back-ticked repo path in DESIGN.md / INTEGRATION.md / README.md / include/b200align.h and in the
comments of the sources must exist, every `test_*` name must be a collected test function."""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DOCS = ["DESIGN.md", "INTEGRATION.md", "README.md", "include/b200align.h"]
SOURCES = (glob.glob(os.path.join(ROOT, "masa-cudalign_b200", "csrc", "*")) + glob.glob(os.path.join(ROOT, "masa-cudalign_b200", "host", "*.[ch]pp")) +
           glob.glob(os.path.join(ROOT, "oracle", "*.c*")) + glob.glob(os.path.join(ROOT, "oracle", "*.sh")) + [os.path.join(ROOT, "bench.py")])
PREFIXES = ("tests/", "profiles/", "tools/", "oracle/", "include/", "masa-cudalign_b200/", "csrc/", "host/")


def _exists(path):
    path = path.rstrip("/.,;:)")
    if any(ch in path for ch in "*{}<>$ "):
        # patterns like profiles/r02_cfg3_n{1,2,4,8}.json or profiles/r02_* : at least one match
        pat = re.sub(r"\{[^}]*\}", "*", path)
        pat = re.sub(r"<[^>]*>", "*", pat)
        return bool(glob.glob(os.path.join(ROOT, pat))) or bool(glob.glob(os.path.join(ROOT, "masa-cudalign_b200", pat)))
    for base in (ROOT, os.path.join(ROOT, "masa-cudalign_b200")):
        if os.path.exists(os.path.join(base, path)):
            return True
    return False


def _cited_paths(text):
    out = set()
    for m in re.finditer(r"[`(\s]((?:%s)[A-Za-z0-9_./{},*<>-]+)" % "|".join(re.escape(p) for p in PREFIXES), text):
        p = m.group(1)
        p = p.split("::")[0].rstrip("/.,;:)")
        if "." not in os.path.basename(p) and "*" not in p:
            continue                                  # prose like "tests/bench only" or a class name, not a file
        if p.endswith(("_ref/", "_ref")) or "/_ref/" in p or p.startswith("oracle/_ref"):
            continue                                  # built artefacts (git-ignored), present only after build()
        out.add(p)
    return out


def _test_names():
    names = set()
    for f in glob.glob(os.path.join(ROOT, "tests", "*.py")):
        names.update(re.findall(r"^def (test_[A-Za-z0-9_]+)", open(f).read(), flags=re.M))
    return names


def test_cited_files_exist():
    missing = []
    for doc in DOCS + SOURCES:
        path = doc if os.path.isabs(doc) else os.path.join(ROOT, doc)
        for p in sorted(_cited_paths(open(path, errors="replace").read())):
            if not _exists(p):
                missing.append(f"{os.path.relpath(path, ROOT)}: {p}")
    assert not missing, "\n".join(missing)


def test_cited_tests_exist():
    have = _test_names()
    missing = []
    for doc in DOCS + SOURCES:
        path = doc if os.path.isabs(doc) else os.path.join(ROOT, doc)
        for name in sorted(set(re.findall(r"\b(test_[a-z0-9_]{6,})\b", open(path, errors="replace").read()))):
            if name.endswith("_cpu") or name.endswith("_gpu") or name == "test_delay_ms":   # file names (checked as paths); a struct field
                continue
            if name not in have:
                missing.append(f"{os.path.relpath(path, ROOT)}: {name}")
    assert not missing, "\n".join(missing)
