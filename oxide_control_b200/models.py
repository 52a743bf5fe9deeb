"""Synthetic inline MJCF models for BASELINE.json's five configs.

All use <compiler angle="radian"/>. They are synthetic stand-ins with the topology and sizes of the
named dm_control / gym tasks, written for this repo (no assets are read from disk at run time).
"""

# C1: single pendulum, 1 hinge, joint damping, no contacts. nq=nv=1, nu=1, nbody=2.
PENDULUM = """
<mujoco model="pendulum">
  <compiler angle="radian"/>
  <option timestep="0.002"/>
  <worldbody>
    <body name="pole" pos="0 0 0">
      <joint name="hinge" type="hinge" axis="0 1 0" damping="0.1"/>
      <geom name="pole" type="capsule" fromto="0 0 0 0 0 -0.5" size="0.02" contype="0" conaffinity="0"/>
    </body>
  </worldbody>
  <actuator>
    <motor name="torque" joint="hinge" gear="1" ctrlrange="-1 1"/>
  </actuator>
  <sensor>
    <jointpos name="angle" joint="hinge"/>
    <jointvel name="anglevel" joint="hinge"/>
  </sensor>
</mujoco>
"""

# C2: cartpole, slide + hinge, no contacts. nq=nv=2, nu=1, nbody=3, nM=3.
CARTPOLE = """
<mujoco model="cartpole">
  <compiler angle="radian"/>
  <option timestep="0.01"/>
  <default>
    <geom contype="0" conaffinity="0"/>
  </default>
  <worldbody>
    <body name="cart" pos="0 0 1">
      <joint name="slider" type="slide" axis="1 0 0" range="-1.8 1.8" damping="0.0005"/>
      <geom name="cart" type="box" size="0.2 0.15 0.1" mass="1"/>
      <body name="pole" pos="0 0 0">
        <joint name="hinge" type="hinge" axis="0 1 0" damping="2e-6"/>
        <geom name="pole" type="capsule" fromto="0 0 0 0 0 1" size="0.045" mass="0.1"/>
      </body>
    </body>
  </worldbody>
  <actuator>
    <motor name="slide" joint="slider" gear="10" ctrlrange="-1 1"/>
  </actuator>
  <sensor>
    <jointpos joint="slider"/>
    <jointpos joint="hinge"/>
    <jointvel joint="slider"/>
    <jointvel joint="hinge"/>
  </sensor>
</mujoco>
"""

# C3: acrobot / double pendulum, RK4, no damping, no contacts. nq=nv=2, nu=1, nbody=3.
ACROBOT = """
<mujoco model="acrobot">
  <compiler angle="radian"/>
  <option timestep="0.01" integrator="RK4"/>
  <default>
    <geom contype="0" conaffinity="0"/>
  </default>
  <worldbody>
    <body name="upper_arm" pos="0 0 2">
      <joint name="shoulder" type="hinge" axis="0 1 0"/>
      <geom name="upper_arm" type="capsule" fromto="0 0 0 0 0 1" size="0.05" mass="1"/>
      <body name="lower_arm" pos="0 0 1">
        <joint name="elbow" type="hinge" axis="0 1 0"/>
        <geom name="lower_arm" type="capsule" fromto="0 0 0 0 0 1" size="0.049" mass="1"/>
      </body>
    </body>
  </worldbody>
  <actuator>
    <motor name="elbow" joint="elbow" gear="2" ctrlrange="-1 1"/>
  </actuator>
</mujoco>
"""

# C4: half-cheetah-style planar chain, 8 capsules + plane, 6 limited+sprung+damped hinges, Newton/pyramidal.
# nq=nv=9, nu=6, nbody=8, ngeom=9, nM=36, ncon<=16, nefc<=12+64.
CHEETAH = """
<mujoco model="cheetah">
  <compiler angle="radian"/>
  <option timestep="0.01"/>
  <default>
    <joint armature="0.1" damping="0.01" limited="true" solimplimit="0 0.8 0.03" solreflimit="0.02 1" stiffness="8"/>
    <geom conaffinity="0" condim="3" contype="1" friction="0.4 0.1 0.1" solimp="0.0 0.8 0.01" solref="0.02 1"/>
    <motor ctrllimited="true" ctrlrange="-1 1"/>
  </default>
  <worldbody>
    <geom name="floor" type="plane" conaffinity="1" pos="0 0 0" size="40 40 40"/>
    <body name="torso" pos="0 0 0.7">
      <joint name="rootx" type="slide" axis="1 0 0" armature="0" damping="0" limited="false" stiffness="0"/>
      <joint name="rootz" type="slide" axis="0 0 1" armature="0" damping="0" limited="false" stiffness="0"/>
      <joint name="rooty" type="hinge" axis="0 1 0" armature="0" damping="0" limited="false" stiffness="0"/>
      <geom name="torso" type="capsule" fromto="-0.5 0 0 0.5 0 0" size="0.046"/>
      <geom name="head" type="capsule" pos="0.6 0 0.1" axisangle="0 1 0 0.87" size="0.046 0.15"/>
      <body name="bthigh" pos="-0.5 0 0">
        <joint name="bthigh" type="hinge" axis="0 1 0" damping="6" range="-0.52 1.05" stiffness="240"/>
        <geom name="bthigh" type="capsule" pos="0.1 0 -0.13" axisangle="0 1 0 -3.8" size="0.046 0.145"/>
        <body name="bshin" pos="0.16 0 -0.25">
          <joint name="bshin" type="hinge" axis="0 1 0" damping="4.5" range="-0.785 0.785" stiffness="180"/>
          <geom name="bshin" type="capsule" pos="-0.14 0 -0.07" axisangle="0 1 0 -2.03" size="0.046 0.15"/>
          <body name="bfoot" pos="-0.28 0 -0.14">
            <joint name="bfoot" type="hinge" axis="0 1 0" damping="3" range="-0.4 0.785" stiffness="120"/>
            <geom name="bfoot" type="capsule" pos="0.03 0 -0.097" axisangle="0 1 0 -0.27" size="0.046 0.094"/>
          </body>
        </body>
      </body>
      <body name="fthigh" pos="0.5 0 0">
        <joint name="fthigh" type="hinge" axis="0 1 0" damping="4.5" range="-1 0.7" stiffness="180"/>
        <geom name="fthigh" type="capsule" pos="-0.07 0 -0.12" axisangle="0 1 0 0.52" size="0.046 0.133"/>
        <body name="fshin" pos="-0.14 0 -0.24">
          <joint name="fshin" type="hinge" axis="0 1 0" damping="3" range="-1.2 0.87" stiffness="120"/>
          <geom name="fshin" type="capsule" pos="0.065 0 -0.09" axisangle="0 1 0 -0.6" size="0.046 0.106"/>
          <body name="ffoot" pos="0.13 0 -0.18">
            <joint name="ffoot" type="hinge" axis="0 1 0" damping="1.5" range="-0.5 0.5" stiffness="60"/>
            <geom name="ffoot" type="capsule" pos="0.045 0 -0.07" axisangle="0 1 0 -0.6" size="0.046 0.07"/>
          </body>
        </body>
      </body>
    </body>
  </worldbody>
  <actuator>
    <motor name="bthigh" joint="bthigh" gear="120"/>
    <motor name="bshin" joint="bshin" gear="90"/>
    <motor name="bfoot" joint="bfoot" gear="60"/>
    <motor name="fthigh" joint="fthigh" gear="120"/>
    <motor name="fshin" joint="fshin" gear="60"/>
    <motor name="ffoot" joint="ffoot" gear="30"/>
  </actuator>
  <sensor>
    <subtreelinvel name="torso_subtreelinvel" body="torso"/>
  </sensor>
</mujoco>
"""

# C5: humanoid-style, free joint + 21 hinges, capsule/sphere geoms vs plane + a filtered set of self pairs.
# nq=28, nv=27, nu=21, nbody=14.
HUMANOID = """
<mujoco model="humanoid">
  <compiler angle="radian"/>
  <option timestep="0.005"/>
  <default>
    <joint type="hinge" limited="true" damping="0.2" stiffness="1" armature="0.01" solimplimit="0 0.99 0.01"/>
    <geom type="capsule" condim="3" friction="0.7 0.005 0.0001" solref="0.015 1" solimp="0.9 0.99 0.003" contype="1" conaffinity="0"/>
    <motor ctrllimited="true" ctrlrange="-1 1"/>
    <default class="big_joint"><joint damping="5" stiffness="10"/></default>
    <default class="big_stiff_joint"><joint damping="5" stiffness="20"/></default>
    <default class="collide_self"><geom contype="3" conaffinity="2"/></default>
  </default>
  <worldbody>
    <geom name="floor" type="plane" conaffinity="1" contype="0" size="100 100 0.2"/>
    <body name="torso" pos="0 0 1.5">
      <freejoint name="root"/>
      <geom name="torso" fromto="0 -0.07 0 0 0.07 0" size="0.07"/>
      <geom name="upper_waist" fromto="-0.01 -0.06 -0.12 -0.01 0.06 -0.12" size="0.06"/>
      <geom name="head" type="sphere" pos="0 0 0.19" size="0.09"/>
      <body name="lower_waist" pos="-0.01 0 -0.26" quat="1 0 -0.002 0">
        <joint name="abdomen_z" axis="0 0 1" pos="0 0 0.065" range="-0.785 0.785" class="big_stiff_joint"/>
        <joint name="abdomen_y" axis="0 1 0" pos="0 0 0.065" range="-1.3 0.52" class="big_joint"/>
        <geom name="lower_waist" fromto="0 -0.06 0 0 0.06 0" size="0.06"/>
        <body name="pelvis" pos="0 0 -0.165" quat="1 0 -0.002 0">
          <joint name="abdomen_x" axis="1 0 0" pos="0 0 0.1" range="-0.61 0.61" class="big_joint"/>
          <geom name="butt" fromto="-0.02 -0.07 0 -0.02 0.07 0" size="0.09"/>
          <body name="right_thigh" pos="0 -0.1 -0.04">
            <joint name="right_hip_x" axis="1 0 0" range="-0.436 0.087" class="big_joint"/>
            <joint name="right_hip_z" axis="0 0 1" range="-1.05 0.61" class="big_joint"/>
            <joint name="right_hip_y" axis="0 1 0" range="-1.92 0.35" class="big_stiff_joint"/>
            <geom name="right_thigh" fromto="0 0 0 0 0.01 -0.34" size="0.06" class="collide_self"/>
            <body name="right_shin" pos="0 0.01 -0.403">
              <joint name="right_knee" axis="0 -1 0" pos="0 0 0.02" range="-2.79 0.035"/>
              <geom name="right_shin" fromto="0 0 0 0 0 -0.3" size="0.049" class="collide_self"/>
              <body name="right_foot" pos="0 0 -0.39">
                <joint name="right_ankle_y" axis="0 1 0" pos="0 0 0.08" range="-0.87 0.87" stiffness="4"/>
                <joint name="right_ankle_x" axis="1 0 0.5" pos="0 0 0.04" range="-0.87 0.87" stiffness="4"/>
                <geom name="right_right_foot" fromto="-0.07 -0.02 0 0.14 -0.04 0" size="0.027" class="collide_self"/>
                <geom name="left_right_foot" fromto="-0.07 0 0 0.14 0.02 0" size="0.027" class="collide_self"/>
              </body>
            </body>
          </body>
          <body name="left_thigh" pos="0 0.1 -0.04">
            <joint name="left_hip_x" axis="-1 0 0" range="-0.436 0.087" class="big_joint"/>
            <joint name="left_hip_z" axis="0 0 -1" range="-1.05 0.61" class="big_joint"/>
            <joint name="left_hip_y" axis="0 1 0" range="-1.92 0.35" class="big_stiff_joint"/>
            <geom name="left_thigh" fromto="0 0 0 0 -0.01 -0.34" size="0.06" class="collide_self"/>
            <body name="left_shin" pos="0 -0.01 -0.403">
              <joint name="left_knee" axis="0 -1 0" pos="0 0 0.02" range="-2.79 0.035"/>
              <geom name="left_shin" fromto="0 0 0 0 0 -0.3" size="0.049" class="collide_self"/>
              <body name="left_foot" pos="0 0 -0.39">
                <joint name="left_ankle_y" axis="0 1 0" pos="0 0 0.08" range="-0.87 0.87" stiffness="4"/>
                <joint name="left_ankle_x" axis="1 0 0.5" pos="0 0 0.04" range="-0.87 0.87" stiffness="4"/>
                <geom name="left_left_foot" fromto="-0.07 0.02 0 0.14 0.04 0" size="0.027" class="collide_self"/>
                <geom name="right_left_foot" fromto="-0.07 0 0 0.14 -0.02 0" size="0.027" class="collide_self"/>
              </body>
            </body>
          </body>
        </body>
      </body>
      <body name="right_upper_arm" pos="0 -0.17 0.06">
        <joint name="right_shoulder1" axis="2 1 1" range="-1.48 1.05"/>
        <joint name="right_shoulder2" axis="0 -1 1" range="-1.48 1.05"/>
        <geom name="right_upper_arm" fromto="0 0 0 0.16 -0.16 -0.16" size="0.04"/>
        <body name="right_lower_arm" pos="0.18 -0.18 -0.18">
          <joint name="right_elbow" axis="0 -1 1" range="-1.57 0.87"/>
          <geom name="right_lower_arm" fromto="0.01 0.01 0.01 0.17 0.17 0.17" size="0.031"/>
          <geom name="right_hand" type="sphere" pos="0.18 0.18 0.18" size="0.04"/>
        </body>
      </body>
      <body name="left_upper_arm" pos="0 0.17 0.06">
        <joint name="left_shoulder1" axis="2 -1 1" range="-1.05 1.48"/>
        <joint name="left_shoulder2" axis="0 1 1" range="-1.05 1.48"/>
        <geom name="left_upper_arm" fromto="0 0 0 0.16 0.16 -0.16" size="0.04"/>
        <body name="left_lower_arm" pos="0.18 0.18 -0.18">
          <joint name="left_elbow" axis="0 -1 -1" range="-1.57 0.87"/>
          <geom name="left_lower_arm" fromto="0.01 -0.01 0.01 0.17 -0.17 0.17" size="0.031"/>
          <geom name="left_hand" type="sphere" pos="0.18 -0.18 0.18" size="0.04"/>
        </body>
      </body>
    </body>
  </worldbody>
  <actuator>
    <motor name="abdomen_y" joint="abdomen_y" gear="40"/>
    <motor name="abdomen_z" joint="abdomen_z" gear="40"/>
    <motor name="abdomen_x" joint="abdomen_x" gear="40"/>
    <motor name="right_hip_x" joint="right_hip_x" gear="40"/>
    <motor name="right_hip_z" joint="right_hip_z" gear="40"/>
    <motor name="right_hip_y" joint="right_hip_y" gear="120"/>
    <motor name="right_knee" joint="right_knee" gear="80"/>
    <motor name="right_ankle_x" joint="right_ankle_x" gear="20"/>
    <motor name="right_ankle_y" joint="right_ankle_y" gear="20"/>
    <motor name="left_hip_x" joint="left_hip_x" gear="40"/>
    <motor name="left_hip_z" joint="left_hip_z" gear="40"/>
    <motor name="left_hip_y" joint="left_hip_y" gear="120"/>
    <motor name="left_knee" joint="left_knee" gear="80"/>
    <motor name="left_ankle_x" joint="left_ankle_x" gear="20"/>
    <motor name="left_ankle_y" joint="left_ankle_y" gear="20"/>
    <motor name="right_shoulder1" joint="right_shoulder1" gear="20"/>
    <motor name="right_shoulder2" joint="right_shoulder2" gear="20"/>
    <motor name="right_elbow" joint="right_elbow" gear="40"/>
    <motor name="left_shoulder1" joint="left_shoulder1" gear="20"/>
    <motor name="left_shoulder2" joint="left_shoulder2" gear="20"/>
    <motor name="left_elbow" joint="left_elbow" gear="40"/>
  </actuator>
</mujoco>
"""

CONFIGS = {
    "pendulum": dict(xml=PENDULUM, nenv=1, precision="f64", label="C1 pendulum"),
    "cartpole": dict(xml=CARTPOLE, nenv=16384, precision="f32", label="C2 cartpole"),
    "acrobot": dict(xml=ACROBOT, nenv=65536, precision="f64", label="C3 acrobot RK4 fp64"),
    "cheetah": dict(xml=CHEETAH, nenv=8192, precision="f32", label="C4 cheetah"),
    "humanoid": dict(xml=HUMANOID, nenv=4096, precision="f32", label="C5 humanoid"),
}
