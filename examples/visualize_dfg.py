"""Print (and optionally draw) the dataflow graph of an algorithm: MFC nodes, the data keys on every edge, and the levels
that may run concurrently.

    python examples/visualize_dfg.py --algo ppo [--dot ppo.dot]        # render with: dot -Tsvg ppo.dot -o ppo.svg

Algorithms: sft, rw, dpo, ppo, grpo, reinforce.  The graphs are taken from the experiment classes themselves (the MFCs a
run would execute), not from hand-written copies, so they cannot drift from the code.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)  # run from a source checkout without installing the package

from realhf_b200.api import dfg  # noqa: E402


def _rpcs(algo: str):
    from realhf_b200.api.quickstart import QUICKSTART_EXPERIMENTS
    import realhf_b200.experiments.algos  # noqa: F401  (registers sft / rw / dpo / ppo / gen)
    if algo in ("grpo", "reinforce") and algo not in QUICKSTART_EXPERIMENTS:  # the example files register themselves on import
        import importlib.util
        spec = importlib.util.spec_from_file_location(f"example_{algo}", os.path.join(ROOT, "examples", "new_algorithms", f"{algo}.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = mod
        spec.loader.exec_module(mod)
    cls = QUICKSTART_EXPERIMENTS[algo]
    exp = cls()
    return list(exp.rpcs.values())


def to_dot(G) -> str:
    lines = ["digraph dfg {", "  rankdir=LR;", '  node [shape=box, style="rounded,filled", fillcolor="#eef3ff", fontname="Helvetica"];']
    for n, d in G.nodes(data=True):
        r = d["object"]
        lines.append(f'  "{n}" [label="{n}\\n{r.interface_type.value} on {r.model_name.role}"];')
    for k in G.graph["dataset_keys"]:
        lines.append(f'  "data:{k}" [label="{k}", shape=ellipse, fillcolor="#f4f4f4"];')
        for c in G.graph["data_consumers"][k]:
            lines.append(f'  "data:{k}" -> "{c}";')
    for u, v, d in G.edges(data=True):
        lines.append(f'  "{u}" -> "{v}" [label="{", ".join(d["keys"])}"];')
    lines.append("}")
    return "\n".join(lines) + "\n"


def main():
    ap = argparse.ArgumentParser("visualize the dataflow graph")
    ap.add_argument("--algo", "-a", default="ppo")
    ap.add_argument("--dot", default=None, help="write a Graphviz file")
    a = ap.parse_args()
    G = dfg.build_graph(_rpcs(a.algo))
    print(f"{a.algo}: {G.number_of_nodes()} model function calls, dataset keys {G.graph['dataset_keys']}")
    for i, level in enumerate(dfg.topological_levels(G)):
        print(f"  level {i}: {', '.join(level)}")
    for u, v, d in G.edges(data=True):
        print(f"  {u} -> {v}: {d['keys']}")
    if a.dot:
        with open(a.dot, "w") as f:
            f.write(to_dot(G))
        print(f"wrote {a.dot}")
    return G


if __name__ == "__main__":
    main()
