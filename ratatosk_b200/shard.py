"""Read sharding across GPUs / ranks: the reference's ticket dealer, made explicit.

The reference hands each worker thread a *ticket*: consecutive reads totalling >= buffer_sz (1 MiB) bases
(src/Ratatosk.cpp:746-772), and restores the input order at the end by ticket id (:919-999).  Multi-GPU
keeps exactly that unit: tickets are dealt round-robin to ranks, every rank corrects its tickets against
its replica of the graph, outputs are merged back in ticket order.  No collective on the data path.
"""

BUFFER_SZ = 1 << 20  # Correct_Opt::buffer_sz (src/Common.hpp:138)


def make_tickets(read_lengths, buffer_sz=BUFFER_SZ):
    """-> list of (first_read, last_read_exclusive): a ticket closes once it holds >= buffer_sz bases"""
    tickets, start, acc = [], 0, 0
    for i, n in enumerate(read_lengths):
        acc += n
        if acc >= buffer_sz:
            tickets.append((start, i + 1))
            start, acc = i + 1, 0
    if start < len(read_lengths):
        tickets.append((start, len(read_lengths)))
    return tickets


def deal(tickets, world_size):
    """round-robin: ticket t goes to rank t % world_size -> per-rank list of ticket ids"""
    return [[t for t in range(len(tickets)) if t % world_size == r] for r in range(world_size)]


def merge_ordered(per_rank_blocks, n_tickets=None):
    """per_rank_blocks: iterable of dict {ticket id: list of output records}; -> records in input order.
    n_tickets (the dealer's count): a missing TRAILING ticket is detected as well"""
    merged = {}
    for blocks in per_rank_blocks:
        for t, recs in blocks.items():
            if t in merged:
                raise ValueError("ticket %d produced twice" % t)
            merged[t] = recs
    out = []
    for t in range(len(merged) if n_tickets is None else n_tickets):
        if t not in merged:
            raise ValueError("ticket %d missing" % t)
        out.extend(merged[t])
    return out
