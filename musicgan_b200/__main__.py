"""Command line with the reference's sub-commands (music_gan/__main__.py:11-120): create_dataset, train, generate."""
import argparse

from . import create_dataset, generate, train


def main() -> None:
    parser = argparse.ArgumentParser("musicgan_b200")
    sub = parser.add_subparsers(dest="mode")
    sub.required = True
    p = sub.add_parser("create_dataset")
    p.add_argument("audio_path", type=str)
    p.add_argument("-o", "--output-dir", type=str, required=True)
    p = sub.add_parser("train")
    p.add_argument("run", type=str)
    p.add_argument("-i", "--input-dataset", dest="input_dataset", type=str, required=True)
    p.add_argument("-o", "--out-path", dest="out_path", type=str, required=True)
    p = sub.add_parser("generate")          # same flags as the reference: positional rand_channels, -n nb-vec, -m nb-music
    p.add_argument("gen_dict_state", type=str)
    p.add_argument("rand_channels", type=int)
    p.add_argument("-n", "--nb-vec", type=int, default=10)
    p.add_argument("-m", "--nb-music", type=int, default=5)
    p.add_argument("-o", "--output-dir", type=str, required=True)
    args = parser.parse_args()
    if args.mode == "create_dataset":
        create_dataset(args.audio_path, args.output_dir)
    elif args.mode == "train":
        train(args.run, args.input_dataset, args.out_path)
    else:
        generate(args.output_dir, args.rand_channels, args.gen_dict_state, args.nb_vec, args.nb_music)


if __name__ == "__main__":
    main()
