"""Box rasteriser with the reference's name (src/utils/bb_to_img.py:5-21): host-side label
preparation (PIL polygon fill), not part of the GPU hot path."""
import numpy as np


def boxes_to_binary_map(x):
    """boxes [N,2,4] (metres; rows x / y; columns fl, fr, bl, br) -> 800x800 float64 ndarray of 0/1:
    pixel = metre * 10 + 400, polygon corners in the order fl, fr, br, bl, image flipped vertically."""
    from PIL import Image, ImageDraw
    x = x.cpu().numpy()
    img = Image.fromarray(np.zeros((800, 800)))
    draw = ImageDraw.Draw(img)
    for i in range(x.shape[0]):
        box = np.stack([x[i][:, 0], x[i][:, 1], x[i][:, 3], x[i][:, 2]]) * 10 + 400
        draw.polygon(list(box.flatten()), fill=1)
    return np.flip(np.asarray(img), 0)
