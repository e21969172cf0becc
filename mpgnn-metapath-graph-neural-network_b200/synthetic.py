"""Seeded, vectorised synthetic-graph generator.

Restates the rules of the reference's `data/synthetic/create_graph_multi_metapath_deterministic.py` -- the script
that produced the `metapath_length_3/4` fixtures and that every configuration beyond the shipped ones depends on --
as array operations, so that the 100k-node (C2) to 10M-node (C4/C5) graphs of BASELINE.json can be made in seconds.
The reference draws one edge at a time from a freshly built N-element candidate list (O(N^2)) and sparsifies with
`DataFrame.iterrows` x list membership; it never seeds its RNG and writes to a hard-coded folder.

Rules kept (reference line numbers):
  * two colours, one-hot node features, colour uniform (:217-226); out-degree uniform in [1, max_rel_for_node] (:228);
  * destination uniform over range(0, N-1) without the source -- node N-1 is never a destination (:240);
  * relation uniform over the preset list of the (source colour, destination colour) pair (:243-244), 16 presets
    indexed by (overlap, shared_relations) (:54-185);
  * the planted relations: one uniform draw per hop from the preset of that hop's colour pair (:196-197);
  * labels by the chained rule (:255-290): stage 0 marks sources of colour c[1] with an edge of relation m[0] into
    colour c[0]; every further stage marks sources with an edge of relation m[i] into a marked node of colour c[i],
    sources restricted to colour c[i+1] except at the last stage (m, c = relations / colours in REVERSED order);
  * sparsification (:368-390): for stage i (forward order) and the nodes marked at that stage, an edge into colour
    c_fwd[i] is deleted iff its relation differs from m_fwd[i], unless some stage marked it as a path edge;
  * the five TSV files (:393-433), byte for byte.
Not kept: `--metapath2/--metapath3`, which cannot run in the reference (`relations_list` is empty at :206).

The deterministic stages are pinned bit for bit by `tests/golden/generator.npz` (recorded from the unmodified
script); the draws use numpy's PCG64 instead of Python's Mersenne twister, so they agree in distribution only
(tested).  Host-side numpy on purpose: this is data preparation in front of K1, not part of the timed path.
"""
import argparse
import os
from dataclasses import dataclass

import numpy as np

COLORS = ("red", "blue")

# (overlap, shared_relations) -> relation ids per colour pair, in the order red-red, red-blue, blue-red, blue-blue
_PRESETS = {
    (0, 0): ([0], [1], [2], [3]),
    (0, 1): ([0, 1], [2, 3], [4, 5], [6, 7]),
    (0, 2): ([0, 1, 2], [3, 4], [5, 6, 7], [8, 9]),
    (0, 3): ([0, 1, 2], [3, 4, 5], [6, 7, 8, 9], [10, 11, 12, 13]),
    (1, 0): ([0, 1], [1], [2, 3], [2]),
    (1, 1): ([0, 7], [1, 2], [2, 3, 5], [3, 4]),
    (1, 2): ([0, 1, 2], [3, 4, 0], [5, 6, 7], [8, 9, 2]),
    (1, 3): ([0, 1, 2, 9], [3, 4, 5, 10], [6, 7, 8, 9], [10, 11, 12, 13]),
    (2, 0): ([0, 3], [1, 2], [2, 3], [0, 1]),
    (2, 1): ([0, 1, 5], [1, 2, 7], [4, 6, 5], [7, 0, 3]),
    (2, 2): ([0, 1, 2, 7], [3, 4, 0], [5, 6, 7], [8, 9, 2, 3]),
    (2, 3): ([0, 1, 2, 9, 8], [3, 4, 5, 10], [6, 7, 8, 9, 11], [10, 11, 12, 13]),
    (3, 0): tuple(list(range(4)) for _ in range(4)),
    (3, 1): tuple(list(range(8)) for _ in range(4)),
    (3, 2): tuple(list(range(10)) for _ in range(4)),
    (3, 3): tuple(list(range(15)) for _ in range(4)),
}


def relation_presets(overlap, shared_relations):
    """Relation ids allowed per (source colour, destination colour), as a [4][...] tuple indexed by 2*src + dst."""
    try:
        return _PRESETS[(int(overlap), int(shared_relations))]
    except KeyError:
        raise ValueError("overlap and shared_relations must be in 0..3 (got %r, %r)" % (overlap, shared_relations))


def _preset_table(presets):
    width = max(len(p) for p in presets)
    table = np.zeros((4, width), dtype=np.int64)
    for i, p in enumerate(presets):
        table[i, :len(p)] = p
    return table, np.array([len(p) for p in presets], dtype=np.int64)


def parse_metapath(metapath):
    """'red-blue-red' -> colour indices in path order; at least two hops (with one the reference's label stage fails)."""
    try:
        cols = [COLORS.index(c) for c in metapath.split("-")]
    except ValueError:
        raise ValueError("metapath colours must be among %s (got %r)" % (COLORS, metapath))
    if len(cols) < 3:
        raise ValueError("the metapath needs at least two hops (three colours): %r" % (metapath,))
    return cols


def plant_metapath(metapath, presets, rng):
    """One relation per hop from the preset of its colour pair.  Returns (relations, colours), both REVERSED -- the
    order the chain is evaluated in and `metapath.dat` stores."""
    cols = parse_metapath(metapath)
    rels = [int(presets[2 * cols[i] + cols[i + 1]][int(rng.integers(0, len(presets[2 * cols[i] + cols[i + 1]])))])
            for i in range(len(cols) - 1)]
    return np.array(rels[::-1], dtype=np.int64), np.array(cols[::-1], dtype=np.int64)


def draw_graph(num_nodes, max_rel_for_node, presets, rng):
    """colours [N] and the (source, relation, destination) list before sparsification, sources ascending."""
    n = int(num_nodes)
    if n < 3:
        raise ValueError("num_nodes must be at least 3")
    if int(max_rel_for_node) < 1:
        raise ValueError("max_rel_for_node must be at least 1")
    colors = rng.integers(0, 2, size=n, dtype=np.int64)
    out_deg = rng.integers(1, int(max_rel_for_node) + 1, size=n, dtype=np.int64)
    src = np.repeat(np.arange(n, dtype=np.int64), out_deg)
    inner = src < n - 1                                   # the last node draws from all of 0..N-2
    k = rng.integers(0, (n - 1) - inner.astype(np.int64))
    dst = k + ((k >= src) & inner)
    table, lengths = _preset_table(presets)
    pair = 2 * colors[src] + colors[dst]
    rel = table[pair, rng.integers(0, lengths[pair])]
    return colors, np.stack([src, rel, dst], axis=1)


def label_chain(colors, triplets, meta_reversed, colors_reversed):
    """Stages of the chained labelling rule, one column per stage in evaluation order; the last column is the label."""
    n, hops = len(colors), len(meta_reversed)
    if hops < 2:
        raise ValueError("the label chain needs at least two hops")
    s, r, d = triplets[:, 0], triplets[:, 1], triplets[:, 2]
    cs, cd = colors[s], colors[d]
    emb = np.zeros((n, hops), dtype=np.int64)
    hit = (cs == colors_reversed[1]) & (r == meta_reversed[0]) & (cd == colors_reversed[0])
    emb[s[hit], 0] = 1
    for i in range(1, hops):
        hit = (r == meta_reversed[i]) & (cd == colors_reversed[i]) & (emb[d, i - 1] == 1)
        if i < hops - 1:
            hit &= cs == colors_reversed[i + 1]
        emb[s[hit], i] = 1
    return emb, emb[:, -1].copy()


def sparsify(colors, triplets, embeddings, meta_reversed, colors_reversed):
    """Boolean keep-mask over the edge list (see the module docstring)."""
    hops = len(meta_reversed)
    meta_fwd, colors_fwd = meta_reversed[::-1], colors_reversed[::-1]
    s, r, d = triplets[:, 0], triplets[:, 1], triplets[:, 2]
    cd = colors[d]
    on_path = np.zeros(len(triplets), dtype=bool)
    competing = np.zeros(len(triplets), dtype=bool)
    for i in range(hops):
        marked = embeddings[:, hops - 1 - i] == 1
        sel = (cd == colors_fwd[i]) & marked[s]
        on_path |= sel & (r == meta_fwd[i])
        competing |= sel & (r != meta_fwd[i])
    return ~(competing & ~on_path)


@dataclass
class SyntheticGraph:
    metapath: str
    colors: np.ndarray            # [N] colour index
    triplets: np.ndarray          # [E, 3] (source, relation, destination), after sparsification
    embeddings: np.ndarray        # [N, hops] stages of the label chain
    labels: np.ndarray            # [N] 0/1
    meta_reversed: np.ndarray     # planted relations, last hop first
    colors_reversed: np.ndarray   # metapath colours, last first
    edges_before_sparsification: int = 0

    @property
    def num_nodes(self):
        return int(len(self.colors))

    @property
    def planted_relations(self):
        """The ground-truth metapath in the order the search reports it and MPNetm consumes it (main.py:1427 prepends,
        layer k reads metapaths[i][k]): index 0 = the hop farthest from the labelled node, last = the relation leaving
        it -- the order of metapath.dat's second line (the fixture's "1 0" is found as [1, 0])."""
        return [int(v) for v in self.meta_reversed]

    def node_features(self):
        x = np.zeros((self.num_nodes, len(COLORS)), dtype=np.int64)
        x[np.arange(self.num_nodes), self.colors] = 1
        return x

    def tensors(self, device=None):
        """(x float32 [N,2], edge_index int64 [2,E], edge_type int64 [E], labels int64 [N]) -- what `load_files` +
        `get_node_features` + `get_edge_index_and_type_no_reverse` return for the written files (main.py:178-195,
        347-372), without the round trip through text."""
        import torch
        x = torch.from_numpy(self.node_features()).float()
        ei = torch.from_numpy(np.ascontiguousarray(self.triplets[:, [0, 2]].T))
        et = torch.from_numpy(np.ascontiguousarray(self.triplets[:, 1]))
        y = torch.from_numpy(self.labels.copy())
        if device is not None:
            x, ei, et, y = x.to(device), ei.to(device), et.to(device), y.to(device)
        return x, ei, et, y

    def write(self, folder):
        """node.dat / link.dat / label.dat / embedding.dat / metapath.dat in the reference's formats."""
        os.makedirs(folder, exist_ok=True)
        ids = np.arange(self.num_nodes, dtype=np.int64)[:, None]
        _write_tsv(os.path.join(folder, "node.dat"), np.hstack([ids, self.node_features()]))
        _write_tsv(os.path.join(folder, "link.dat"), self.triplets)
        _write_tsv(os.path.join(folder, "label.dat"), np.hstack([ids, self.labels[:, None]]))
        _write_tsv(os.path.join(folder, "embedding.dat"), np.hstack([ids, self.embeddings]), trailing_tab=True)
        with open(os.path.join(folder, "metapath.dat"), "w") as f:
            f.write(self.metapath + "\n")
            f.write("".join("%d " % v for v in self.meta_reversed) + "\n")
            f.write("".join("%d " % v for v in self.colors_reversed))


def _write_tsv(path, table, trailing_tab=False):
    """Integer matrix -> tab-separated lines ('a\tb\tc\n'; embedding.dat ends every line with a tab, as the reference
    writes it).  pyarrow's multi-threaded CSV writer when it is installed (10M-node graphs: seconds instead of
    minutes), numpy.savetxt otherwise -- same bytes either way (tested)."""
    table = np.ascontiguousarray(table, dtype=np.int64)
    try:
        import pyarrow as pa
        import pyarrow.csv as pacsv
    except ImportError:
        fmt = "\t".join(["%d"] * table.shape[1]) + ("\t" if trailing_tab else "")
        np.savetxt(path, table, fmt=fmt)
        return
    cols = [pa.array(table[:, j]) for j in range(table.shape[1])]
    if trailing_tab:
        cols.append(pa.repeat(pa.scalar(""), len(table)))     # an empty last field = the trailing tab
    names = ["c%d" % j for j in range(len(cols))]
    pacsv.write_csv(pa.Table.from_arrays(cols, names=names), path,
                    write_options=pacsv.WriteOptions(include_header=False, delimiter="\t", quoting_style="none"))


def disjoint_presets(num_relations):
    """num_relations / 4 relations per colour pair, no relation shared between pairs: the reference's (overlap 0,
    shared 0) preset `([0], [1], [2], [3])` scaled up (configs with more relations than the 16 presets hold)."""
    if num_relations < 4 or num_relations % 4 != 0:
        raise ValueError("num_relations must be a positive multiple of 4")
    q = num_relations // 4
    return tuple(list(range(i * q, (i + 1) * q)) for i in range(4))


def generate(num_nodes, max_rel_for_node, metapath, overlap, shared_relations, seed=0, sparsification=True, presets=None):
    """The reference script's `main(args)` with a seed: draw the graph, plant the metapath, label, sparsify.
    `presets` (four relation-id lists, one per colour pair) replaces the (overlap, shared_relations) table entry."""
    presets = relation_presets(overlap, shared_relations) if presets is None else presets
    rng = np.random.Generator(np.random.PCG64(int(seed)))
    meta_rev, cols_rev = plant_metapath(metapath, presets, rng)        # drawn first, as in the reference (:196)
    colors, triplets = draw_graph(num_nodes, max_rel_for_node, presets, rng)
    emb, labels = label_chain(colors, triplets, meta_rev, cols_rev)
    before = len(triplets)
    if sparsification:
        triplets = triplets[sparsify(colors, triplets, emb, meta_rev, cols_rev)]
    return SyntheticGraph(metapath, colors, triplets, emb, labels, meta_rev, cols_rev, before)


def main(argv=None):
    ap = argparse.ArgumentParser(description="synthetic graph creation (seeded, vectorised)")
    ap.add_argument("--num_nodes", type=int, required=True, help="number of nodes")
    ap.add_argument("--max_rel_for_node", type=int, required=True, help="maximum number of outgoing edges for node")
    ap.add_argument("--metapath", type=str, required=True, help="target metapath, e.g. red-blue-red-blue")
    ap.add_argument("--overlap", type=int, required=True, help="relations overlap (0..3)")
    ap.add_argument("--shared_relations", type=int, required=True, help="shared_relations (0..3)")
    ap.add_argument("--metapath2", type=str, default=None, help="not supported (cannot run in the reference either)")
    ap.add_argument("--metapath3", type=str, default=None, help="not supported (ignored by the reference)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--out", type=str, default=None,
                    help="output folder (default: ./overlap_<o>rels_<s>, the reference's folder name)")
    args = ap.parse_args(argv)
    if args.metapath2 or args.metapath3:
        raise NotImplementedError("--metapath2/--metapath3: the reference fails at its line 206 with a second metapath")
    g = generate(args.num_nodes, args.max_rel_for_node, args.metapath, args.overlap, args.shared_relations, args.seed)
    out = args.out or "overlap_%drels_%d" % (args.overlap, args.shared_relations)
    g.write(out)
    print("%s: %d nodes, %d edges (%d before sparsification), %d positive, planted relations %s" %
          (out, g.num_nodes, len(g.triplets), g.edges_before_sparsification, int(g.labels.sum()), g.planted_relations))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
