"""Label tables of the editing path (data): the 19 face-parsing classes produced by the BiSeNet
parser and the 40 CelebA attributes of the AnyCost-GAN predictor, in the reference's index order
(src/constants.py)."""

_FACE_PARTS = ("background skin l_brow r_brow l_eye r_eye eye_g l_ear r_ear ear_r nose mouth u_lip "
               "l_lip neck neck_l cloth hair hat")
ATTRS = _FACE_PARTS.split()
ATTR_DICT = {name: i for i, name in enumerate(ATTRS)}

_CELEBA = ("5_o_Clock_Shadow Arched_Eyebrows Attractive Bags_Under_Eyes Bald Bangs Big_Lips Big_Nose "
           "Black_Hair Blond_Hair Blurry Brown_Hair Bushy_Eyebrows Chubby Double_Chin Eyeglasses Goatee "
           "Gray_Hair Heavy_Makeup High_Cheekbones Male Mouth_Slightly_Open Mustache Narrow_Eyes No_Beard "
           "Oval_Face Pale_Skin Pointy_Nose Receding_Hairline Rosy_Cheeks Sideburns Smiling Straight_Hair "
           "Wavy_Hair Wearing_Earrings Wearing_Hat Wearing_Lipstick Wearing_Necklace Wearing_Necktie Young")
ANY_GAN_ATTRS = _CELEBA.split()
ANY_GAN_ATTRS_DICT = {name: i for i, name in enumerate(ANY_GAN_ATTRS)}

assert len(ATTRS) == 19 and len(ANY_GAN_ATTRS) == 40
