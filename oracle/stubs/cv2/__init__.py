"""Stub: OpenCV is only used by the reference's debug viewer (tsc/legged_gym/envs/base/legged_robot.py:287-296)."""
WINDOW_NORMAL = 0
COLORMAP_JET = 2


def _no(*a, **k):
    raise NotImplementedError("cv2 stub")


namedWindow = imshow = waitKey = applyColorMap = convertScaleAbs = _no
