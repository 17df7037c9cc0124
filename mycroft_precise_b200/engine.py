"""``precise-engine`` wire protocol on the GPU (reference: precise/scripts/engine.py:15-73).

    python -m mycroft_precise_b200.engine MODEL [CHUNK_SIZE] < audio.raw

stdin: raw int16 mono 16 kHz audio written in groups of CHUNK_SIZE *bytes*; stdout: one ASCII float per
chunk, flushed after every line; everything else goes to stderr; -v/--version prints the version;
a tty stdin is refused; EOF / KeyboardInterrupt end the loop quietly.  CHUNK_SIZE = -1 reads until EOF
and prints a single prediction, like the reference.  This is what
``precise_runner.PreciseEngine(exe_file, model_file, chunk_size)`` spawns (runner.py:50-55).
"""
import argparse
import sys


def main(argv=None):
    from . import __version__
    ap = argparse.ArgumentParser(prog='precise-engine', description=__doc__,
                                 formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument('-v', '--version', action='version', version=__version__)
    ap.add_argument('model_name', help='weights file (.npz, .net or .pb) with its .params next to it')
    ap.add_argument('chunk_size', type=int, nargs='?', default=-1,
                    help='Number of bytes to read before making a prediction')
    ap.add_argument('--device', type=int, default=0)
    args = ap.parse_args(argv)
    if sys.stdin.isatty():
        raise ValueError('Please pipe audio via stdin using < audio.wav')

    stdout = sys.stdout
    sys.stdout = sys.stderr                      # keep the wire clean (engine.py:55-56)
    try:
        from .runner import B200Listener, B200Engine, _resolve_model
        from .params import ListenerParams
        pr = _resolve_model(args.model_name)[1] or ListenerParams()
        # The stateful device tick (B200Engine) covers even chunk sizes that complete at most 8 frames per tick; anything else
        # (odd sizes, very large chunks, CHUNK_SIZE = -1) goes through the stateless mirror of Listener.update.
        samples = args.chunk_size // 2
        if args.chunk_size > 0 and args.chunk_size % 2 == 0 and samples // pr.hop_samples + 2 <= 8:
            eng = B200Engine(args.model_name, args.chunk_size, device=args.device)
            eng.start()
            # Audio the window still depends on, cut at a multiple of hop_samples from the start of the stream: a short last
            # read (the reference's Listener.update still answers it, network_runner.py:132-153) is replayed through the mirror.
            hist = bytearray()
            keep = 2 * (pr.buffer_samples + pr.window_samples + pr.hop_samples)
            state = {'dropped': 0}

            def step():
                chunk = sys.stdin.buffer.read(args.chunk_size)
                if len(chunk) == 0:
                    raise EOFError
                if len(chunk) < args.chunk_size:           # ragged last read: same answer as the reference, through the mirror
                    lis = B200Listener(args.model_name, -1, device=args.device)
                    if len(hist):
                        lis.update(bytes(hist))
                    return lis.update(chunk[:len(chunk) & ~1])
                hist.extend(chunk)
                if len(hist) > 2 * keep:
                    cut = len(hist) - keep
                    cut -= (state['dropped'] + cut) % (2 * pr.hop_samples)     # keep the frame grid aligned
                    if cut > 0:
                        del hist[:cut]
                        state['dropped'] += cut
                return eng.get_prediction(chunk)
        else:
            lis = B200Listener(args.model_name, args.chunk_size, device=args.device)

            def step():
                return lis.update(sys.stdin.buffer)
        try:
            while True:
                conf = step()
                stdout.buffer.write((str(conf) + '\n').encode('ascii'))
                stdout.buffer.flush()
        except (EOFError, KeyboardInterrupt):
            pass
    finally:
        sys.stdout = stdout


if __name__ == '__main__':
    main()
