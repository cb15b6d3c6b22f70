"""Host side of read generation (mirrors src/esloco/read_operations.py).  The placement loop itself runs in the
CUDA library; here are the small host computations that feed it and the dict-shaped view of its result."""
import numpy as np

LABEL_NAMES = {-1: "Blocked/Consuming", -2: "Blocked/NotConsuming"}


def get_weighted_probabilities(insertion_name, n_barcodes, weights_dict):
    """Share of one barcode (read_operations.py:69-82): a weight key applies when it equals any "_"-separated
    part of the barcode name (or the whole name); unweighted barcodes count 1; README.md:204-210."""
    if weights_dict is None:
        return 1 / n_barcodes
    common_denominator = (sum(weights_dict.values()) + n_barcodes - len(weights_dict)) * n_barcodes
    parts = insertion_name.split("_")
    for key in weights_dict:
        if key in parts or key == insertion_name:
            return (weights_dict[key] * n_barcodes) / common_denominator
    return n_barcodes / common_denominator


def barcode_probabilities(n_barcodes, barcode_weights):
    """Normalised barcode probabilities (read_operations.py:94-97)."""
    p = [get_weighted_probabilities(f"Barcode_{i}", n_barcodes, barcode_weights) for i in range(n_barcodes)]
    return p / np.sum(p)


def barcode_cdf(probabilities):
    """The CDF numpy's choice(names, p=) searches with side='right' (read_operations.py:101)."""
    cdf = np.array(probabilities, dtype=np.float64).cumsum()
    cdf /= cdf[-1]
    return cdf


def generate_read_length_distribution(num_reads, mean_read_length, min_read_length, sigma, distribution="lognormal",
                                      rng=None):
    """Length pool (read_operations.py:8-27): round(lognormal(ln(mean) - 0.5, sigma)) kept when > min_read_length.
    Drawn on the host with a numpy Generator (set-up, once per combination)."""
    rng = rng or np.random.default_rng()
    if distribution == "normal":
        x = rng.normal(mean_read_length, mean_read_length / 10, num_reads)
    elif distribution == "lognormal":
        x = rng.lognormal(mean=np.log(mean_read_length) - 0.5, sigma=sigma, size=num_reads)
    else:
        raise ValueError("Unsupported distribution. Supported options: 'lognormal'.")
    x = np.round(x).astype(np.int64)
    return x[x > min_read_length]


def reads_to_dict(start, length, label):
    """Reference-shaped read dict: "Read_{n}_{label}" -> (start, end) in draw order (read_operations.py:113-119)."""
    out = {}
    for n, (s, ln, lb) in enumerate(zip(start.tolist(), length.tolist(), label.tolist())):
        name = LABEL_NAMES.get(lb) or f"Barcode_{lb}"
        out[f"Read_{n}_{name}"] = (s, s + ln)
    return out


def dict_to_reads(read_dir):
    """Inverse of reads_to_dict: arrays (start, length, label) from a reference-shaped read dict; the label is the
    last "_" token of the key (counting.py:37,66)."""
    n = len(read_dir)
    start = np.empty(n, np.int64); length = np.empty(n, np.int64); label = np.empty(n, np.int32)
    for i, (k, (s, e)) in enumerate(read_dir.items()):
        tok = k.split("_")[-1]
        if tok == "Blocked/Consuming":
            label[i] = -1
        elif tok == "Blocked/NotConsuming":
            label[i] = -2
        else:
            label[i] = int(tok)
        start[i] = int(s); length[i] = int(e) - int(s)
    return start, length, label


def generate_reads_based_on_coverage(genome_size, coverage, read_length_distribution, n_barcodes, barcode_weights, consuming,
                                     masked_tree, seed=0, iteration=0, combo=0, device=0):
    """generate_reads_based_on_coverage with the reference's signature (read_operations.py:86-123) -> (read dict
    "Read_{n}_{label}" -> (start, end) in draw order, covered_length), placed by the CUDA engine and copied back.
    Meant for inspection and small runs: the simulation itself never materialises reads.  seed / iteration / combo
    key the Philox stream."""
    from .session import get_session
    sess = get_session(device)
    sess.invalidate()
    # the caller hands over the very lengths to pick from (read_operations.py:102): exact uniform picks, not the quantile table
    sess.configure(genome_size, {}, masked_tree, n_barcodes, barcode_weights, length_sampler="pool")
    lengths = np.asarray(read_length_distribution)
    sess.set_pool(None, lambda: lengths)
    eng = sess.engine
    res = eng.run_batch(seed, combo, iteration, 1, coverage, consuming).cpu()
    if int(res.status[0]) & 2:
        raise RuntimeError("coverage not reached")
    start, length, label = (x.cpu().numpy() for x in eng.materialize_reads(seed, combo, iteration, consuming, int(res.n_reads[0])))
    return reads_to_dict(start, length, label), int(res.n_bases[0])
