"""Extract the numeric task constants of the reference's MJCF files into factory_marl_b200/assets.json.

Run in the build container only (reads /root/reference, which does not exist on the GPU box):
    python tools/extract_assets.py
Sources: challenge_env/challenge_env/assets/kuka_iiwa_14/iiwa14.xml (bodies :55-147, excludes :150-158,
defaults :7-30, actuators :160-168) and assets/gripper.xml (bodies :4-44, equality :46-49, excludes :50-54,
tendon :55-60, actuator :61-65).  Only numbers/structure are extracted; no code.
"""
import json
import os
import sys
import xml.etree.ElementTree as ET

REF = "/root/reference/challenge_env/challenge_env/assets"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "factory_marl_b200", "assets.json")


def fl(s):
    return [float(x) for x in s.split()]


def parse_defaults(node, inherited, table):
    cur = {k: dict(v) for k, v in inherited.items()}
    for ch in node:
        if ch.tag == "default":
            continue
        cur.setdefault(ch.tag, {}).update(ch.attrib)
    if "class" in node.attrib:
        table[node.attrib["class"]] = cur
    for ch in node:
        if ch.tag == "default":
            parse_defaults(ch, cur, table)


def parse_body(b, classes, childclass):
    childclass = b.attrib.get("childclass", childclass)
    nd = {"name": b.attrib["name"]}
    if "pos" in b.attrib:
        nd["pos"] = fl(b.attrib["pos"])
    if "quat" in b.attrib:
        nd["quat"] = fl(b.attrib["quat"])
    ine = b.find("inertial")
    if ine is not None:
        d = {"mass": float(ine.attrib["mass"]), "pos": fl(ine.attrib.get("pos", "0 0 0")),
             "diaginertia": fl(ine.attrib["diaginertia"])}
        if "quat" in ine.attrib:
            d["quat"] = fl(ine.attrib["quat"])
        nd["inertial"] = d
    j = b.find("joint")
    if j is not None:
        cls = classes.get(j.attrib.get("class", childclass), {}).get("joint", {})
        at = dict(cls)
        at.update(j.attrib)
        jt = {"type": at.get("type", "hinge"), "axis": fl(at.get("axis", "0 0 1"))}
        if "range" in at:
            jt["range"] = fl(at["range"])
        nd["joint"] = jt
    geoms = []
    for g in b.findall("geom"):
        cls = classes.get(g.attrib.get("class", childclass), {}).get("geom", {})
        at = dict(cls)
        at.update(g.attrib)
        t = at.get("type", "sphere")
        if t == "mesh":
            assert at.get("contype") == "0" and at.get("conaffinity") == "0"
            geoms.append({"type": "mesh"})
        elif t == "sphere":
            geoms.append({"type": "sphere", "size": float(at["size"]), "pos": fl(at.get("pos", "0 0 0"))})
        elif t == "box":
            geoms.append({"type": "box", "size": fl(at["size"]), "pos": fl(at.get("pos", "0 0 0")),
                          "friction": fl(at["friction"]), "solimp": fl(at["solimp"]), "solref": fl(at["solref"]),
                          "priority": int(at["priority"])})
        else:
            raise ValueError(t)
    nd["geoms"] = geoms
    nd["children"] = [parse_body(c, classes, childclass) for c in b.findall("body")]
    return nd


def main():
    iiwa = ET.parse(f"{REF}/kuka_iiwa_14/iiwa14.xml").getroot()
    classes = {}
    for d in iiwa.find("default"):
        if d.tag == "default":
            parse_defaults(d, {}, classes)
    gen = classes["iiwa"]["general"]
    out = {"iiwa": {
        "bodies": [parse_body(iiwa.find("worldbody").find("body"), classes, None)],
        "excludes": [[e.attrib["body1"], e.attrib["body2"]] for e in iiwa.find("contact")],
        "gainprm": fl(gen["gainprm"])[0], "biasprm": fl(gen["biasprm"]),
    }}
    # attachment site
    for s in iiwa.iter("site"):
        if s.attrib.get("name") == "attachment_site":
            out["attachment_site_pos"] = fl(s.attrib["pos"])
    gr = ET.parse(f"{REF}/gripper.xml").getroot()
    act = gr.find("actuator").find("general")
    eq = gr.find("equality").find("joint")
    site = [s for s in gr.iter("site")][0]
    out["gripper"] = {
        "bodies": [parse_body(gr.find("worldbody").find("body"), {}, None)],
        "excludes": [[e.attrib["body1"], e.attrib["body2"]] for e in gr.find("contact")],
        "gainprm": fl(act.attrib["gainprm"])[0], "biasprm": fl(act.attrib["biasprm"]),
        "ctrlrange": fl(act.attrib["ctrlrange"]), "forcerange": fl(act.attrib["forcerange"]),
        "tendon_coef": [float(j.attrib["coef"]) for j in gr.find("tendon").find("fixed")],
        "eq_solref": fl(eq.attrib["solref"]), "eq_solimp": fl(eq.attrib["solimp"]),
        "site_pos": fl(site.attrib["pos"]),
    }
    # checks on structure the model compiler relies on
    assert [j.attrib["joint"] for j in gr.find("tendon").find("fixed")] == ["left_plate_slide_joint", "right_plate_slide_joint"]
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    sys.exit(main())
