"""Documentation hygiene: every relative link and every `path`-looking reference to a repository file in the guides resolves."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DOCS = sorted([os.path.join("docs", f) for f in os.listdir(os.path.join(ROOT, "docs")) if f.endswith(".md")]
              + ["README.md", "DESIGN.md", "PARITY.md", "ROADMAP.md", os.path.join("profiles", "README.md")])


@pytest.mark.parametrize("doc", DOCS)
def test_markdown_links_resolve(doc):
    text = open(os.path.join(ROOT, doc)).read()
    base = os.path.dirname(os.path.join(ROOT, doc))
    for target in re.findall(r"\]\(([^)#\s]+)(?:#[^)]*)?\)", text):
        if re.match(r"^[a-z]+://", target):
            continue
        assert os.path.exists(os.path.normpath(os.path.join(base, target))), f"{doc}: broken link {target}"


@pytest.mark.parametrize("doc", DOCS)
def test_backticked_repository_paths_exist(doc):
    """`realhf_b200/...`, `tests/...`, `scripts/...`, `examples/...`, `docs/...`, `profiles/...` in backticks must exist (globs
    and brace lists are expanded loosely; `path::symbol` checks the path part)."""
    text = open(os.path.join(ROOT, doc)).read()
    missing = []
    for ref in set(re.findall(r"`((?:realhf_b200|tests|scripts|examples|docs|profiles)/[A-Za-z0-9_./*{},\-]+)", text)):
        path = ref.split("::")[0].rstrip(".,")
        if path.startswith("docs/source"):   # citations of the reference's documentation tree
            continue
        if any(ch in path for ch in "*{"):
            path = re.split(r"[*{]", path)[0].rstrip("/")
            path = os.path.dirname(path) if not os.path.isdir(os.path.join(ROOT, path)) else path
        if path and not os.path.exists(os.path.join(ROOT, path)):
            missing.append(ref)
    assert not missing, f"{doc}: references to files that do not exist: {sorted(missing)}"
