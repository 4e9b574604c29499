"""Import-safe shim: the webcam front end (reference scanner/acquisition/camera.py) is outside
the hot path.  `from scanner.acquisition import Camera` (src/3-capture_decode.py:5) imports; the
class opens a cv2.VideoCapture only when instantiated."""


class Camera(object):
    def __init__(self, src=0, width=1920, height=1080, fps=30):
        import cv2
        self._cap = cv2.VideoCapture(src)
        self._cap.set(cv2.CAP_PROP_FRAME_WIDTH, width)
        self._cap.set(cv2.CAP_PROP_FRAME_HEIGHT, height)
        self._cap.set(cv2.CAP_PROP_FPS, fps)

    def get_frame(self):
        ok, frame = self._cap.read()
        return frame if ok else None

    def stop_cam(self):
        self._cap.release()
