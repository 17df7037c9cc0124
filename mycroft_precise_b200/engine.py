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
        from .runner import B200Listener, B200Engine
        if args.chunk_size > 0 and args.chunk_size % 2 == 0:
            eng = B200Engine(args.model_name, args.chunk_size, device=args.device)
            eng.start()

            def step():
                chunk = sys.stdin.buffer.read(args.chunk_size)
                if len(chunk) < args.chunk_size:           # b'' or a ragged last read: the reference's
                    raise EOFError                         # Listener would block/raise here as well
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
