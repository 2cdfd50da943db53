"""Worker process of the clustering pool (``python -m clustertracking_b200._find_worker``).

Reads pickled tasks from stdin, labels its frames straight into shared memory and answers with the
per-frame label spans (or the exception) on stdout.  Host-side only: no CUDA, no torch."""
import pickle
import sys


def main():
    from clustertracking_b200 import find
    inp, out = sys.stdin.buffer, sys.stdout.buffer
    while True:
        try:
            task = pickle.load(inp)
        except EOFError:
            return
        try:
            reply = find._pool_task(task)
        except Exception as exc:          # sent back and re-raised in the parent
            reply = exc
        pickle.dump(reply, out)
        out.flush()


if __name__ == "__main__":
    main()
