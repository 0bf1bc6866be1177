"""The signal-pad handshake of csrc/exchange.cu (``meet_peers``), model-checked on the CPU.

Every (rank, peer) pair is served by one thread of block b that, per barrier, flips ITS word in the peer's pad 0 -> 1
(spinning while the word still holds the previous round's 1) and then flips the PEER's word in its own pad 1 -> 0.  The
words are shared between the two barriers of a launch and between launches.  This test explores EVERY interleaving of
those compare-and-swap steps for small worlds and checks what the kernel relies on:
  * no deadlock: some thread can always move until all have finished all rounds,
  * barrier: a thread leaves round k only after its peer's thread has entered round k (so the peer's earlier writes --
    its gradients before the first barrier, its parameter stores before the second -- are ordered before),
  * the pads return to all-zero, so the next launch starts from the state the first one found.
It checks the protocol, not the CUDA code: the GPU run of tools/exchange_check.py does that."""
import itertools


def explore(world, rounds):
    pairs = [(a, b) for a in range(world) for b in range(world)]          # thread (a, b): rank a's thread for peer b
    # thread state: 2 * round + phase (0 = about to put, 1 = about to wait); done at 2 * rounds
    start = (tuple(0 for _ in pairs), tuple(0 for _ in pairs))           # (thread states, pad words); word index = pairs.index((owner, sender))
    word = {p: i for i, p in enumerate(pairs)}                           # (owner rank, sender rank) -> index
    seen, stack, finals = {start}, [start], 0
    while stack:
        threads, pads = stack.pop()
        moved = False
        for ti, (a, b) in enumerate(pairs):
            st = threads[ti]
            if st == 2 * rounds:
                continue
            k, phase = divmod(st, 2)
            if phase == 0:                                               # CAS(peer's pad[me], 0 -> 1)
                w = word[(b, a)]
                if pads[w] != 0:
                    continue
                new_pads = pads[:w] + (1,) + pads[w + 1:]
            else:                                                        # CAS(my pad[peer], 1 -> 0)
                w = word[(a, b)]
                if pads[w] != 1:
                    continue
                # the signal consumed here was put by thread (b, a) in round k: it has entered round k
                assert threads[pairs.index((b, a))] >= 2 * k + 1, "left a barrier before the peer arrived"
                new_pads = pads[:w] + (0,) + pads[w + 1:]
            moved = True
            nxt = (threads[:ti] + (st + 1,) + threads[ti + 1:], new_pads)
            if nxt not in seen:
                seen.add(nxt)
                stack.append(nxt)
        if not moved:
            assert all(s == 2 * rounds for s in threads), f"deadlock at {threads} {pads}"
            assert not any(pads), "pads must return to zero"
            finals += 1
    return len(seen), finals


def test_handshake_has_no_deadlock_and_is_a_barrier():
    for world, rounds in ((1, 4), (2, 2), (2, 4), (3, 2)):               # 2 barriers per launch: 4 rounds = two launches
        states, finals = explore(world, rounds)
        assert finals == 1 and states > 1


def test_signal_words_of_different_blocks_and_ranks_are_disjoint():
    """Word layout of the kernel: signal_base + block * world + sender, inside the owner's pad."""
    from jaxngp_b200 import exchange as X
    for world in (2, 4, 8):
        n_blocks = X.blocks_for(9216, world)
        words = [X.SIGNAL_BASE + b * world + s for b, s in itertools.product(range(n_blocks), range(world))]
        assert len(set(words)) == len(words) and min(words) >= X.SIGNAL_BASE and (max(words) + 1) * 4 <= 9216
