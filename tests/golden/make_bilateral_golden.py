#!/usr/bin/env python
"""Golden vectors for the ingest oracle (oracle/oracle_ingest.cpp), generated in the development
container with OpenCV 4.13's CPU functions -- the CPU twins of the cv::cuda functions the reference
calls (core/src/supersurfel_fusion.cu:175,180; their CUDA sources are an un-vendored dependency):

  cv2.bilateralFilter(depth, -1, 0.03, 4.5)   same radius / disc / reflect-101 rules as
      cv::cuda::bilateralFilter, colour weights through an interpolated table -> agrees with
      the exact-weight restatement to ~1e-6 m
  Mat::convertTo(CV_32F, scale) of a 16-bit depth image (one fp32 product per pixel)

(cv2's RGB2GRAY uses 15-bit coefficients where OpenCV 3.4's CUDA path uses the 14-bit ones, so
it differs from the restatement by at most one grey level; it is stored only to document that.)

  python tests/golden/make_bilateral_golden.py     # writes tests/golden/ingest_golden.npz
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from supersurfel_fusion_b200.synth import SyntheticSequence  # noqa: E402


def main():
    seq = SyntheticSequence(width=160, height=120, seed=4321)
    rgb, depth = seq.frame(3)
    depth16 = np.round(depth * 5000.0).astype(np.uint16)            # TUM encoding, 0 = missing
    scale = np.float32(0.0002)
    # Mat::convertTo(CV_32F, scale) of 16-bit data is one fp32 product per pixel (cv2 has no direct binding)
    decoded = (depth16.astype(np.float32) * scale).astype(np.float32)
    filtered = cv2.bilateralFilter(decoded, -1, 0.03, 4.5)
    gray = cv2.cvtColor(rgb, cv2.COLOR_RGB2GRAY)
    np.savez_compressed(os.path.join(HERE, "ingest_golden.npz"), rgb=rgb, depth16=depth16, scale=scale,
                        decoded=decoded, filtered=filtered, gray_cv2=gray, cv2_version=cv2.__version__)
    print("wrote ingest_golden.npz", filtered.shape, cv2.__version__)


if __name__ == "__main__":
    main()
