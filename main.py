"""CLI of the B200-native video-feature engine: the flags of the reference's main.py (main.py:93-149), same names,
defaults and choices.  ``--device_ids`` starts one process per GPU (reference: one thread per GPU); ``--cpu`` is
refused -- the reference's CPU path is what ``bench.py --impl reference`` times, the engine itself has no CPU path.
"""
import argparse
import functools

import numpy  # noqa: F401  (kept first, as in the reference)
import torch  # noqa: F401

from video_features_b200.utils import form_list_from_user_input, sanity_check

SUPPORTED = ['i3d', 'raft', 'CLIP-ViT-B/32', 'CLIP4CLIP-ViT-B-32']


def build_extractor(args):
    """feature_type -> extractor (main.py:15-41)."""
    if args.feature_type in ['CLIP-ViT-B/32', 'CLIP4CLIP-ViT-B-32']:
        from video_features_b200.extract.extract_clip import ExtractCLIP
        return ExtractCLIP(args)
    if args.feature_type == 'i3d':
        from video_features_b200.extract.extract_i3d import ExtractI3D
        return ExtractI3D(args)
    if args.feature_type == 'raft':
        from video_features_b200.extract.extract_raft import ExtractRAFT
        return ExtractRAFT(args)
    if args.feature_type in ['vggish', 'r21d_rgb', 'resnet18', 'resnet34', 'resnet50', 'resnet101', 'resnet152', 'pwc',
                             'CLIP-ViT-B/16', 'vggish_torch']:
        raise NotImplementedError(f'{args.feature_type}: outside the hot path this engine rebuilds (SURVEY.md §2)')
    raise NotADirectoryError                      # main.py:41


def parallel_feature_extraction(args):
    from video_features_b200.dispatch import parallel_feature_extraction as run
    video_paths = form_list_from_user_input(args)
    run(functools.partial(build_extractor, args), len(video_paths), args.device_ids)


def make_parser():
    parser = argparse.ArgumentParser(description='Extract Features')
    parser.add_argument('--feature_type', required=True,
                        choices=['i3d', 'vggish', 'r21d_rgb', 'resnet18', 'resnet34', 'resnet50', 'resnet101',
                                 'resnet152', 'raft', 'pwc', 'CLIP-ViT-B/32', 'CLIP-ViT-B/16', 'CLIP4CLIP-ViT-B-32',
                                 'vggish_torch'])
    parser.add_argument('--video_paths', nargs='+', help='space-separated paths to videos')
    parser.add_argument('--flow_paths', nargs='+', help='space-separated paths to video flow images')
    parser.add_argument('--file_with_video_paths', help='.txt file where each line is a path')
    parser.add_argument('--video_dir', type=str, help='dir of videos')
    parser.add_argument('--flow_dir', type=str,
                        help='dir of optical flow of videos. [flow_dir]/[video id]/[flow_(x/y)_000001.jpg]')
    parser.add_argument('--device_ids', type=int, nargs='+', help='space-separated device ids')
    parser.add_argument('--cpu', action='store_true', help='use cpu only')
    parser.add_argument('--tmp_path', default='./tmp',
                        help='folder to store the temporary files used for extraction (frames or aud files)')
    parser.add_argument('--keep_tmp_files', dest='keep_tmp_files', action='store_true', default=False,
                        help='to keep temp files after feature extraction. (works only for vggish and i3d)')
    parser.add_argument('--on_extraction', default='print', choices=['print', 'save_numpy', 'save_pickle'],
                        help='what to do once the stack is extracted')
    parser.add_argument('--output_path', default='./output', help='where to store results if saved')
    parser.add_argument('--output_direct', action="store_true",
                        help='if so, files will be directly saved in output_path')
    parser.add_argument('--extraction_fps', type=float, help='(Outdated)For original video fps, leave unspecified')
    parser.add_argument('--extract_method', type=str, help='extraction frames method.')
    parser.add_argument('--stack_size', type=int, help='Feature time span in fps')
    parser.add_argument('--step_size', type=int, help='Feature step size in fps')
    parser.add_argument('--streams', nargs='+', choices=['flow', 'rgb'],
                        help='Streams to use for feature extraction. Both used if not specified')
    parser.add_argument('--flow_type', choices=['raft', 'pwc', 'flow'], default='pwc',
                        help='Flow to use in I3D. PWC is faster while RAFT is more accurate.')
    parser.add_argument('--batch_size', type=int, default=1,
                        help='Batchsize (only frame-wise extractors are supported)')
    parser.add_argument('--resize_to_larger_edge', dest='resize_to_smaller_edge', action='store_false',
                        default=True, help='The larger side will be resized to this number maintaining the'
                        + 'aspect ratio. By default, uses the smaller side (as Resize in torchvision).')
    parser.add_argument('--side_size', type=int,
                        help='If specified, the input images will be resized to this value in RAFT.')
    parser.add_argument('--show_pred', dest='show_pred', action='store_true', default=False,
                        help='to show preds of a model, i.e. on a pre-train dataset (imagenet or kinetics) for each feature')
    return parser


if __name__ == "__main__":
    args = make_parser().parse_args()
    if args.on_extraction in ['save_numpy', 'save_pickle']:
        print(f'Saving features to {args.output_path}')
    if args.keep_tmp_files:
        print(f'Keeping temp files in {args.tmp_path}')
    sanity_check(args)
    if args.cpu:
        raise SystemExit('--cpu: this engine has no CPU path (the reference CPU flow is timed by '
                         '`python bench.py --impl reference`); pass --device_ids')
    if not args.device_ids:
        args.device_ids = [0]
    parallel_feature_extraction(args)
