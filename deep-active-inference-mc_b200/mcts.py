"""Host-side MCTS planner over the EFE evaluators — the caller side of the hot path.

Same entry point, parameter bag and return tuple as the reference's src/mcts.py
(`active_inference_mcts(model, frame, params, o_shape)` :150-195, `MCTS_Params` :137-148), and the same
search semantics, so a decision made here equals the decision src/mcts.py makes when both drive the same
model under the same noise (tests/test_planner.py checks that against the real reference where it is
mounted).  It is not a copy: the tree is a structure of arrays (one row per node) instead of linked Node
objects, which is what a batched-leaf / device-resident variant needs next (§8 f rank 2).

Per expansion the model is called exactly like the reference does:
    expand    -> model.calculate_G_mean(s x4, eye(4))            (use_means)       src/mcts.py:78
                 model.calculate_G(s x4, eye(4), samples=S)      (otherwise)       src/mcts.py:81
    simulate  -> model.mcts_step_simulate(s, depth, use_means=False)               src/mcts.py:188
`params.samples` (optional, default 1 like Node.expand's default) is the one addition: BASELINE.json's
config 4 needs N=50 samples per expansion and src/mcts.py never passes its `samples` argument.
"""
import torch


class MCTS_Params:
    """src/mcts.py:137-148, plus the optional `samples`."""

    def __init__(self):
        self.C = 1.0
        self.threshold = 0.5
        self.repeats = 300
        self.simulation_repeats = 1
        self.simulation_depth = 3
        self.use_habit = False
        self.use_means = True
        self.verbose = False
        self.method = 'ai'
        self.using_prior_for_exploration = False
        self.samples = 1


def calc_threshold(P, axis):
    return torch.max(P, dim=axis).values - torch.mean(P, dim=axis)


def normalization(x, tau=1):
    return x / x.sum(dim=0)


class Tree:
    """Search tree as arrays: node i has state[i] (latent), W[i], N[i] (per action), child[i] (node ids or -1)."""

    def __init__(self, pi_dim, capacity, C, use_prior):
        self.pi_dim, self.C, self.use_prior = pi_dim, C, use_prior
        self.W = torch.zeros(capacity, pi_dim)
        self.N = torch.zeros(capacity, pi_dim)
        self.Qpi = torch.zeros(capacity, pi_dim)
        self.child = torch.full((capacity, pi_dim), -1, dtype=torch.long)
        self.state = [None] * capacity
        self.size = 0

    def add(self, s):
        i = self.size
        self.state[i] = s
        self.size += 1
        return i

    def is_leaf(self, i):
        return bool((self.child[i] < 0).any())

    def selection_scores(self, i):
        """src/mcts.py:39-47: Q normalised to a distribution plus the C/N exploration bonus."""
        q = self.W[i] / self.N[i]
        q = q - q.min()
        q = q / q.sum()
        if self.use_prior:
            return q + self.C * self.Qpi[i] * 1.0 / self.N[i]
        return q + self.C * 1.0 / self.N[i]

    def select(self, root):
        """Descend by argmax score until a leaf (src/mcts.py:49-62); returns (node path below root, actions)."""
        nodes, actions, cur = [], [], root
        while True:
            a = int(torch.argmax(self.selection_scores(cur)))
            actions.append(a)
            cur = int(self.child[cur, a])
            nodes.append(cur)
            if self.is_leaf(cur):
                return nodes, actions

    def expand(self, i, model, use_means, samples):
        """src/mcts.py:64-86: one 4-row EFE evaluation; W -= G, N += 1, four children from the next states."""
        s4 = torch.stack([self.state[i]] * self.pi_dim)
        pi_hot = model.pi_one_hot if self.pi_dim == 4 else model.pi_one_hot_3
        if use_means:
            G, _, nxt, _ = model.calculate_G_mean(s4, pi_hot)
        else:
            G, _, nxt, _, _ = model.calculate_G(s4, pi_hot, samples=samples)
        self.W[i] -= G.detach().to("cpu")
        self.N[i] += 1.0
        for a in range(self.pi_dim):
            self.child[i, a] = self.add(nxt[a])

    # ---- batched leaves -------------------------------------------------------------------
    def select_batch(self, root, k):
        """Up to k DISTINCT leaves for one batched expansion.  Descents are made one after the other with the
        reference's argmax rule (selection_scores); after each, the edges of its path get one virtual visit (N + 1 with
        the edge's mean W/N kept, so only the C/N exploration bonus shrinks) and the claimed leaf is blocked — as is
        any node whose children are all blocked — so the next descent goes elsewhere.  With k = 1 nothing virtual is
        ever read: the selection is the reference's.  Returns [(nodes, actions)], statistics restored."""
        picks, blocked = [], set()
        W0, N0 = self.W.clone(), self.N.clone()
        parent = {}
        while len(picks) < k and root not in blocked:
            nodes, actions, cur = [], [], root
            while True:
                sc = self.selection_scores(cur).clone()
                for a in range(self.pi_dim):
                    if int(self.child[cur, a]) in blocked:
                        sc[a] = -float("inf")
                a = int(torch.argmax(sc))
                nxt = int(self.child[cur, a])
                parent[nxt] = cur
                actions.append(a)
                nodes.append(nxt)
                cur = nxt
                if self.is_leaf(cur):
                    break
            picks.append((nodes, actions))
            blocked.add(cur)
            up = cur
            while up != root:                       # a node with no unblocked child is exhausted for this batch
                up = parent[up]
                if all(int(c) in blocked for c in self.child[up]):
                    blocked.add(up)
                else:
                    break
            for i, a in zip([root] + nodes[:-1], actions):
                q = self.W[i, a] / self.N[i, a]
                self.N[i, a] += 1.0
                self.W[i, a] = q * self.N[i, a]
        self.W, self.N = W0, N0
        return picks

    def expand_batch(self, leaves, model, use_means, samples):
        """expand() for several leaves with ONE model call: rows = leaf*4 + action (the calculate_G_repeated layout,
        src/util.py:57-60)."""
        pi_hot = model.pi_one_hot if self.pi_dim == 4 else model.pi_one_hot_3
        s = torch.stack([self.state[i] for i in leaves for _ in range(self.pi_dim)])
        pi = torch.as_tensor(pi_hot).repeat(len(leaves), 1)
        if use_means:
            G, _, nxt, _ = model.calculate_G_mean(s, pi)
        else:
            G, _, nxt, _, _ = model.calculate_G(s, pi, samples=samples)
        G = G.detach().to("cpu").reshape(len(leaves), self.pi_dim)
        for j, i in enumerate(leaves):
            self.W[i] -= G[j]
            self.N[i] += 1.0
            for a in range(self.pi_dim):
                self.child[i, a] = self.add(nxt[j * self.pi_dim + a])

    def backpropagate(self, nodes, actions, G):
        for i, a in zip(nodes, actions):
            self.W[i, a] -= G
            self.N[i, a] += 1

    def most_visited_path(self, root):
        """src/mcts.py:98-127: follow argmax N to a leaf, then drop opposite-action pairs."""
        path, cur = [], root
        while True:
            a = int(torch.argmax(self.N[cur]))
            path.append(a)
            cur = int(self.child[cur, a])
            if self.is_leaf(cur):
                break
        return trim_path(path, self.pi_dim)


def trim_path(path, pi_dim):
    """src/mcts.py:108-126: drop pairs of opposite actions (and, as there, the last action of the path)."""
    if pi_dim == 4:
        opposite = {(0, 1), (1, 0), (2, 3), (3, 2)}
    elif pi_dim == 3:
        opposite = {(1, 2), (2, 1)}
    else:
        raise ValueError(f'Error: Unknown number of pi_dim {pi_dim}')
    out, i = [], 0
    while i < len(path) - 1:
        if (path[i], path[i + 1]) in opposite:
            i += 2
        else:
            out.append(path[i])
            i += 1
    return out


def active_inference_mcts(model, frame, params, o_shape=(64, 64, 1)):
    """src/mcts.py:150-195.  Returns (path, repeats_done, states_explored, all_paths, all_paths_G)."""
    states_explored, all_paths, all_paths_G = 0, [], []
    if frame is None or (hasattr(frame, "__len__") and len(frame) == 0):
        return [0], 0, states_explored, all_paths, all_paths_G
    samples = int(getattr(params, "samples", 1))
    frame = torch.as_tensor(frame)
    qs0_mean, _ = model.model_down.encoder(frame.reshape(1, *o_shape))
    tree = Tree(model.pi_dim, 1 + model.pi_dim * (params.repeats + 2), params.C, params.using_prior_for_exploration)
    root = tree.add(qs0_mean[0])
    root_qpi = model.model_top.encode_s(qs0_mean)[1][0]
    tree.Qpi[root] = root_qpi.detach().to("cpu")
    if params.use_habit and calc_threshold(tree.Qpi[root], axis=0) > params.threshold:
        return [torch.multinomial(tree.Qpi[root], 1).item()], 0, states_explored, all_paths, all_paths_G

    tree.expand(root, model, params.use_means, samples)
    for repeat in range(params.repeats):
        if calc_threshold(normalization(tree.N[root]), axis=0) > params.threshold:
            return tree.most_visited_path(root), repeat, states_explored, all_paths, all_paths_G
        nodes, actions = tree.select(root)
        leaf = nodes[-1]
        tree.expand(leaf, model, params.use_means, samples)
        sims = torch.zeros(params.simulation_repeats)
        for k in range(params.simulation_repeats):
            states_explored += params.simulation_depth
            sims[k], _, qpi = model.mcts_step_simulate(tree.state[leaf], params.simulation_depth, use_means=False)
            tree.Qpi[leaf] = qpi.detach().to("cpu")
        tree.backpropagate([root] + nodes[:-1], actions, sims.mean())
        all_paths.append(actions)
        all_paths_G.append(sims.mean().item())
    return tree.most_visited_path(root), params.repeats, states_explored, all_paths, all_paths_G


def active_inference_mcts_batched(model, frame, params, o_shape=(64, 64, 1), leaves=8):
    """Batched-leaf variant of active_inference_mcts: every iteration claims up to `leaves` distinct
    leaves (Tree.select_batch), expands them with ONE EFE evaluation of 4*leaves rows and simulates them with ONE
    model.mcts_step_simulate_batch pass — two host waits per `leaves` expansions instead of two per expansion, and
    batches large enough to occupy the GPU.  params.repeats still counts expansions.  leaves = 1 makes the decisions of
    active_inference_mcts call for call (same model calls, same noise)."""
    states_explored, all_paths, all_paths_G = 0, [], []
    if frame is None or (hasattr(frame, "__len__") and len(frame) == 0):
        return [0], 0, states_explored, all_paths, all_paths_G
    samples = int(getattr(params, "samples", 1))
    frame = torch.as_tensor(frame)
    qs0_mean, _ = model.model_down.encoder(frame.reshape(1, *o_shape))
    tree = Tree(model.pi_dim, 1 + model.pi_dim * (params.repeats + leaves + 2), params.C, params.using_prior_for_exploration)
    root = tree.add(qs0_mean[0])
    tree.Qpi[root] = model.model_top.encode_s(qs0_mean)[1][0].detach().to("cpu")
    if params.use_habit and calc_threshold(tree.Qpi[root], axis=0) > params.threshold:
        return [torch.multinomial(tree.Qpi[root], 1).item()], 0, states_explored, all_paths, all_paths_G

    tree.expand(root, model, params.use_means, samples)
    done = 0
    while done < params.repeats:
        if calc_threshold(normalization(tree.N[root]), axis=0) > params.threshold:
            return tree.most_visited_path(root), done, states_explored, all_paths, all_paths_G
        picks = tree.select_batch(root, min(leaves, params.repeats - done))
        ids = [nodes[-1] for nodes, _ in picks]
        starts = torch.stack([tree.state[i] for i in ids])          # read before the expansion appends children
        tree.expand_batch(ids, model, params.use_means, samples)
        sims = torch.zeros(len(picks))
        for _ in range(params.simulation_repeats):
            states_explored += params.simulation_depth * len(picks)
            G, _, qpi = model.mcts_step_simulate_batch(starts, params.simulation_depth, use_means=False)
            sims += torch.as_tensor(G, dtype=torch.float32).reshape(-1)
            tree.Qpi[ids] = qpi.detach().to("cpu")
        sims /= params.simulation_repeats
        for (nodes, actions), g in zip(picks, sims):
            tree.backpropagate([root] + nodes[:-1], actions, g)
            all_paths.append(actions)
            all_paths_G.append(g.item())
        done += len(picks)
    return tree.most_visited_path(root), done, states_explored, all_paths, all_paths_G


def active_inference_mcts_device(model, frame, params, o_shape=(64, 64, 1), leaves=8):
    """active_inference_mcts_batched with the search tree resident on the GPU (dai_mcts_plan): selection, expansion
    bookkeeping, back-propagation and the final path are kernels between the EFE evaluations and the host waits once
    per decision.  Same return tuple and — same model calls under the same noise keys — the same decisions as
    active_inference_mcts_batched(model, ..., leaves) (tests/test_planner.py); `model` must be the CUDA model."""
    if frame is None or (hasattr(frame, "__len__") and len(frame) == 0):
        return [0], 0, 0, [], []
    frame = torch.as_tensor(frame)
    qs0_mean = None
    if params.use_habit:                                  # src/mcts.py:166-170: needs the prior on the host first
        qs0_mean, _ = model.model_down.encoder(frame.reshape(1, *o_shape))
        qpi = model.model_top.encode_s(qs0_mean)[1][0].detach().to("cpu")
        if calc_threshold(qpi, axis=0) > params.threshold:
            return [torch.multinomial(qpi, 1).item()], 0, 0, [], []
    model._sync()
    raw, done, _, all_paths, all_G = model._engine.mcts_plan(None if qs0_mean is not None else frame, params, leaves,
                                                             qs0_mean=None if qs0_mean is None else qs0_mean[0])
    explored = done * params.simulation_depth * params.simulation_repeats
    return trim_path(raw, model.pi_dim), done, explored, all_paths, all_G
