"""CPU oracle of the milliEye detection-and-fusion hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under millieye_b200/ imports this package; the only callers
are tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, where
it is the checker or the timed CPU baseline, never the product.

What it restates (reference = sxontheway/milliEye @ 8d23df9, paths under
module3_our_dataset/):
  parse_config.py   utils/parse_config.py:3-21
  darknet.py        yolov3/models.py:12-79 (module construction rules), :132-179 (YOLO decode),
                    :247-267 (forward) - conv/BN/pool arithmetic is delegated to the same torch
                    CPU operators the reference's nn.Modules call (fp32)
  boxes.py          utils/utils.py:58-74 (box converts), :337-378 (non_max_suppression_cpp) and the
                    third-party torchvision ops it calls: batched_nms (ops/boxes.py:51-120 of
                    torchvision 0.26.0, the version installed in this image; reference pinned
                    0.6.0, README.md:11) and the greedy nms kernel, restated in numpy
  roi.py            torchvision ps_roi_align / roi_align (third-party, same version note),
                    restated in numpy with the semantics probed in SURVEY.md §8a A10/A11
  fusion.py         my_models.py:433-539 (Network.forward, inference branch) and the heads
                    :47-77, :130-157, :176-210, :213-284, :378-391
  stage3_loss.py    my_models.py:545-640 (labelling + loss branch: obtain_iou_labels :317-375, FocalLoss :287-314,
                    regression_loss :394-408, sampling, metric)
  stage3_train.py   one train-mode forward + backward of the fusion heads as train.py:169-186 runs it (torch autograd):
                    the checker of the backward pass the product does not have yet
  stage3_backward.py the same step's backward derived by hand, op by op (BatchNorm on batch statistics, conv wgrad /
                    dgrad, RoIAlign adjoint, heads, losses) - the arithmetic the next round's kernels implement;
                    checked against stage3_train.py
  yolo_loss.py      yolov3/models.py:180-232 + utils/utils.py:381-440 (YOLO training loss of Darknet.forward(x, targets);
                    checker only - the product does not implement that branch)
  radar.py          utils/datasets.py:56-106 (radar heat-map) and data_collection/utils/utils.py:81-120 (projection)
  synth.py          deterministic synthetic weights (numpy RandomState) shared by the golden
                    generator and the tests

Pinning: the reference has no golden vectors or tests.  The oracle is pinned
against (a) outputs of the reference itself, imported read-only in the build container by
tests/golden/make_golden.py and committed as tests/golden/*.npz, and (b) the installed
torchvision 0.26.0 CPU ops for the third-party pieces (tests/test_oracle_*.py).
"""
